"""north_star boundary clause: "Python host code hands tensors to CUDA via DLPack through a thin C-ABI layer (ctypes ...),
with no PyTorch".  The package must import and solve with PyTorch absent: numpy in, numpy / DLPack out, device
memory and copies from the library's own host surface (csrc/xde_hostapi.cu, paddlexde_b200/_native.py)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BLOCK = "import sys\nsys.modules['torch'] = None  # any `import torch` now raises ImportError\n"


def run_without_torch(body: str) -> str:
    r = subprocess.run([sys.executable, "-c", BLOCK + textwrap.dedent(body)], cwd=ROOT, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_package_imports_without_torch():
    out = run_without_torch("""
        import paddlexde_b200 as px
        assert px._tensor.torch is None
        for name in ("odeint", "odeint_adjoint", "sdeint", "ddeint", "Dopri5", "RK4", "Euler", "MLPField"):
            assert hasattr(px, name), name
        from paddlexde_b200._native import DeviceArray
        from paddlexde_b200.functional.odeint_adjoint import adjoint_backward
        try:
            px.odeint_adjoint(None, None, None, solver=px.Dopri5)
        except ImportError as e:
            assert "autograd" in str(e)
        else:
            raise SystemExit("odeint_adjoint must ask for an autograd framework")
        assert "torch" not in [m for m in sys.modules if sys.modules[m] is not None]
        print("ok")
    """)
    assert out.strip().endswith("ok")


@pytest.mark.gpu
def test_cfg1_through_odeint_without_torch():
    """cfg1 (spiral ODE, MLP 2-50-2, dopri5, batch 20) and the adjoint, an RK4 solve, an SDE solve and a history gather in
    a process that cannot import torch: ctypes + numpy only, results bit-exact against the oracle."""
    out = run_without_torch("""
        import numpy as np
        import paddlexde_b200 as px
        from oracle import xde_oracle as xo
        from paddlexde_b200._native import DeviceArray
        from paddlexde_b200.functional.odeint_adjoint import adjoint_backward
        from tests.problems import cfg2_tspan, fanin_weights, spiral_weights

        w = spiral_weights()
        field, om = px.MLPField(*w, pre="cube"), xo.MLP(*w, pre="cube")
        assert isinstance(field.w1, DeviceArray)
        rng = np.random.default_rng(42)
        y0 = (np.array([2.0, 0.0]) + rng.standard_normal((20, 2))).astype(np.float32)
        t = cfg2_tspan(32)
        sol = px.odeint(field, y0, t, px.Dopri5, options={"controller": "trajectory"})
        ref, st, _, rc = xo.dopri5_mlp(om, y0, t)
        assert isinstance(sol, np.ndarray) and rc == 0 and np.array_equal(sol, ref)
        # device-resident in, device-resident out, DLPack / CUDA array interface on the result
        sol_d = px.odeint(field, DeviceArray.from_numpy(y0), t, px.Dopri5, options={"controller": "trajectory"})
        assert isinstance(sol_d, DeviceArray) and np.array_equal(sol_d.numpy(), ref)
        assert sol_d.__dlpack_device__() == (2, 0) and sol_d.__cuda_array_interface__["shape"] == ref.shape
        cap = sol_d.__dlpack__()
        assert "dltensor" in repr(cap)
        # the adjoint as a plain function
        gy = np.zeros_like(ref); gy[-1] = np.sign(ref[-1]) / ref[-1].size
        g, a0, stats, _ = adjoint_backward(field, t, ref, gy, return_adj_y0=True)
        g_ref, a_ref, st_ref, _, rc = xo.dopri5_mlp_adjoint(om, t, ref, gy)
        assert rc == 0 and np.array_equal(a0.numpy(), a_ref)
        assert np.allclose(g.numpy(), g_ref, rtol=1e-5, atol=1e-6 * np.abs(g_ref).max())
        assert stats.read().n_attempts == int(st_ref.n_attempts.sum())
        # fixed grid (FP32 kernels and the tcgen05 path), both y0 layouts
        out = px.odeint(field, y0, t[:6], px.RK4)
        assert np.array_equal(out, xo.fixed_mlp("rk4", om, y0, t[:6]).transpose(1, 0, 2).reshape(-1, 2))
        wt = fanin_weights(64, 256, seed=1)
        yt = rng.uniform(-1, 1, (200, 1, 64)).astype(np.float32)
        tt = np.linspace(0, 1, 6).astype(np.float32)
        tens = px.odeint(px.MLPField(*wt, pre="id"), yt, tt, px.RK4, options={"math": "tensor"})
        ref_t = xo.fixed_mlp("rk4", xo.MLP(*wt, "id"), yt[:, 0], tt)
        assert np.allclose(tens, ref_t, rtol=1e-5, atol=1e-5 * np.abs(ref_t).max())
        # SDE with supplied increments, history gather
        f4, g4 = fanin_weights(4, 32, seed=2), fanin_weights(4, 32, seed=3)
        ys = rng.uniform(-1, 1, (33, 1, 4)).astype(np.float32)
        dW = (0.2 * rng.standard_normal((5, 33, 4))).astype(np.float32)
        sde = px.sdeint(px.MLPField(*f4, pre="cube"), px.MLPField(*g4, pre="square"), ys, tt, px.Euler,
                        options={"bm_increments": dW})
        assert np.array_equal(sde, xo.sde_mlp("em", xo.MLP(*f4, "cube"), xo.MLP(*g4, "square"), ys[:, 0], tt, dW))
        his = rng.uniform(-1, 1, (3, 7, 40, 3)).astype(np.float32)
        span = np.arange(40, dtype=np.float32)
        lags = np.array([0.5, 3.25, 17.0, 38.9], np.float32)
        val, der = px.xde.base_dde.history_gather(lags, his, span, "cubic")
        v_ref, d_ref = xo.history_gather("cubic", his, span, lags)
        assert np.array_equal(val.numpy(), v_ref) and np.array_equal(der.numpy(), d_ref)
        assert px.launch_count() >= 7
        print("ok")
    """)
    assert out.strip().endswith("ok")


@pytest.mark.gpu
def test_device_array_interchange_with_torch():
    """DLPack producer of the C ABI (xde_dlpack_wrap): torch imports a DeviceArray without a copy, and the solvers take
    DeviceArray inputs inside a PyTorch process."""
    import torch

    import paddlexde_b200 as px
    from paddlexde_b200._native import DeviceArray

    a = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    d = DeviceArray.from_numpy(a)
    t = torch.from_dlpack(d)
    assert t.is_cuda and t.data_ptr() == d.data_ptr() and torch.equal(t.cpu(), torch.from_numpy(a))
    t.mul_(2)
    assert np.array_equal(d.numpy(), 2 * a)  # same memory
    t2 = torch.as_tensor(d, device="cuda")   # __cuda_array_interface__
    assert t2.data_ptr() == d.data_ptr()
    assert np.array_equal(d[1].numpy(), 2 * a[1]) and np.array_equal(d.reshape(6, 4)[2:5].numpy(), 2 * a.reshape(6, 4)[2:5])
