"""Import the reference's hot-path modules -- its UNMODIFIED source files under /root/reference -- on the NumPy `paddle`
stand-in, without executing the package `__init__`s (those pull in the whole product: CDE, Brownian trees, scipy
wrappers, third-party packages that are not installed).  TEST INFRASTRUCTURE ONLY; needs /root/reference, so it is
used by tools/make_reference_golden.py and by the CPU tests that re-derive the committed vectors when the reference
tree is present (never on the GPU box)."""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

REF = os.environ.get("XDE_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "paddlexde", "solver"))


def _pkg(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def load():
    """-> namespace with the reference's classes / functions (executed from its own files)."""
    if not available():
        raise FileNotFoundError(f"{REF}/paddlexde not found")
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)  # `import paddle` -> oracle/ref_shim/paddle
    saved = {k: v for k, v in sys.modules.items() if k == "paddle" or k.startswith("paddle.")}
    for k in saved:
        del sys.modules[k]
    import paddle  # noqa: F401  (the stand-in)

    assert os.path.dirname(paddle.__file__).startswith(_HERE), "a real paddle shadows the stand-in"
    root = os.path.join(REF, "paddlexde")
    for k in [k for k in sys.modules if k == "paddlexde" or k.startswith("paddlexde.")]:
        del sys.modules[k]
    _pkg("paddlexde", root)
    xde = _pkg("paddlexde.xde", os.path.join(root, "xde"))
    base_xde = _load("paddlexde.xde.base_xde", os.path.join(root, "xde", "base_xde.py"))
    xde.BaseXDE = base_xde.BaseXDE
    base_ode = _load("paddlexde.xde.base_ode", os.path.join(root, "xde", "base_ode.py"))
    xde.BaseODE = base_ode.BaseODE
    utils = _pkg("paddlexde.utils", os.path.join(root, "utils"))
    ode_utils = _load("paddlexde.utils.ode_utils", os.path.join(root, "utils", "ode_utils.py"))
    utils.ode_utils = ode_utils
    interp = _pkg("paddlexde.interpolation", os.path.join(root, "interpolation"))
    _pkg("paddlexde.interpolation.functional", os.path.join(root, "interpolation", "functional"))
    interp_fn = _load("paddlexde.interpolation.functional.interp_fn",
                      os.path.join(root, "interpolation", "functional", "interp_fn.py"))
    sys.modules["paddlexde.interpolation.functional"].linear_interp = interp_fn.linear_interp
    sys.modules["paddlexde.interpolation.functional"].cubic_hermite_interp = interp_fn.cubic_hermite_interp
    interp.functional = sys.modules["paddlexde.interpolation.functional"]
    _load("paddlexde.interpolation.interpolate_base", os.path.join(root, "interpolation", "interpolate_base.py"))
    interp.interpolate = _load("paddlexde.interpolation.interpolate", os.path.join(root, "interpolation", "interpolate.py"))
    base_dde = _load("paddlexde.xde.base_dde", os.path.join(root, "xde", "base_dde.py"))
    xde.BaseDDE = base_dde.BaseDDE
    _pkg("paddlexde.solver", os.path.join(root, "solver"))
    _load("paddlexde.solver.base_adaptive_solver", os.path.join(root, "solver", "base_adaptive_solver.py"))
    rk = _load("paddlexde.solver.base_adaptive_solver_rk", os.path.join(root, "solver", "base_adaptive_solver_rk.py"))
    fixed = _load("paddlexde.solver.base_fixed_solver", os.path.join(root, "solver", "base_fixed_solver.py"))
    _pkg("paddlexde.solver.adaptive_solver", os.path.join(root, "solver", "adaptive_solver"))
    _pkg("paddlexde.solver.fixed_solver", os.path.join(root, "solver", "fixed_solver"))
    utils.misc = _load("paddlexde.utils.misc", os.path.join(root, "utils", "misc.py"))
    _pkg("paddlexde.functional", os.path.join(root, "functional"))
    f_odeint = _load("paddlexde.functional.odeint", os.path.join(root, "functional", "odeint.py"))
    f_adjoint = _load("paddlexde.functional.odeint_adjoint", os.path.join(root, "functional", "odeint_adjoint.py"))
    f_ddeint = _load("paddlexde.functional.ddeint", os.path.join(root, "functional", "ddeint.py"))
    ns = types.SimpleNamespace(paddle=paddle, ode_utils=ode_utils, BaseODE=base_ode.BaseODE, rk=rk, fixed=fixed,
                               interp_fn=interp_fn, odeint_mod=f_odeint, adjoint_mod=f_adjoint, interpolate=interp.interpolate,
                               base_dde=base_dde, ddeint=f_ddeint.ddeint)
    for mod, cls in (("dopri5", "Dopri5"), ("bosh3", "Bosh3"), ("fehlberg2", "Fehlberg2"), ("adaptive_heun", "AdaptiveHeun"),
                     ("dopri8", "Dopri8")):
        m = _load(f"paddlexde.solver.adaptive_solver.{mod}", os.path.join(root, "solver", "adaptive_solver", mod + ".py"))
        setattr(ns, cls, getattr(m, cls))
    for mod, cls in (("euler", "Euler"), ("midpoint", "Midpoint"), ("rk4", "RK4")):
        m = _load(f"paddlexde.solver.fixed_solver.{mod}", os.path.join(root, "solver", "fixed_solver", mod + ".py"))
        setattr(ns, cls, getattr(m, cls))
    return ns
