"""Development check of the tcgen05 path (csrc/xde_tc.cu): tensor vs the bit-exact FP32 kernels on the
device, error statistics and timings per shape.  Usage: python tools/tc_check.py [quick]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import paddlexde_b200._lib as _L
if os.environ.get("XDE_LIB"):  # kernel tuning experiments: an alternative build of the library
    _L._SO = os.path.abspath(os.environ["XDE_LIB"])
import paddlexde_b200 as px
from tests.problems import fanin_weights


def timeit(fn, n=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def err(a, b):
    a, b = a.double(), b.double()
    scale = b.abs().max().item()
    return {"max_abs": (a - b).abs().max().item(), "max_abs_over_scale": (a - b).abs().max().item() / scale,
            "max_rel_1e-3floor": ((a - b).abs() / b.abs().clamp_min(1e-3 * scale)).max().item(),
            "finite": bool(torch.isfinite(a).all().item())}


def f64_field(w, pre):
    w1, b1, w2, b2 = [torch.from_numpy(np.asarray(a)).cuda().double() for a in w]
    p = {"id": 1, "square": 2, "cube": 3}[pre]
    return lambda y: torch.tanh((y ** p) @ w1 + b1) @ w2 + b2


def f64_fixed(f, y0, t, solver):
    """fp64 restatement of Euler.step / rk4_alt_step_func (base_fixed_solver.py:166-197)."""
    y = y0.double()
    for i in range(1, len(t)):
        dt = float(t[i]) - float(t[i - 1])
        k1 = f(y)
        if solver == "Euler":
            y = y + dt * k1
        else:
            k2 = f(y + dt / 3 * k1)
            k3 = f(y + dt * (k1 - k2 / 3))
            k4 = f(y + dt * (k1 - k2 + k3))
            y = y + dt * (k1 + 3 * k2 + 3 * k3 + k4) / 8
    return y


def ode(d, h, B, solver, pre, steps=10, time=False):
    w = fanin_weights(d, h, seed=d + h)
    field = px.MLPField(*w, pre=pre)
    y0 = torch.from_numpy(np.random.default_rng(d).uniform(-1, 1, (B, 1, d)).astype(np.float32)).cuda()
    t = np.linspace(0, 1, steps + 1).astype(np.float32)
    S = getattr(px, solver)
    a = px.odeint(field, y0, t, S, options={"math": "tensor"})
    b = px.odeint(field, y0, t, S, options={"math": "fp32"})
    r = {"case": f"ode {solver} {d}-{h}-{d} pre={pre} B={B} steps={steps}", **err(a, b)}
    if B <= 4096:
        ref = f64_fixed(f64_field(w, pre), y0.reshape(B, d), t, solver)
        et, ef = err(a[:, -1], ref), err(b[:, -1], ref)
        r.update({"tensor_vs_f64": et["max_abs_over_scale"], "fp32_vs_f64": ef["max_abs_over_scale"],
                  "ratio": et["max_abs_over_scale"] / max(ef["max_abs_over_scale"], 1e-30)})
    if time:
        r["ms_tensor"] = timeit(lambda: px.odeint(field, y0, t, S, options={"math": "tensor", "out_stride": steps, "check_status": False}))
        r["ms_fp32"] = timeit(lambda: px.odeint(field, y0, t, S, options={"math": "fp32", "out_stride": steps}))
    return r


def sde(d, h, B, time=False):
    f = px.MLPField(*fanin_weights(d, h, seed=2), pre="cube")
    g = px.MLPField(*fanin_weights(d, h, seed=3), pre="square")
    gen = torch.Generator(device="cuda").manual_seed(2)
    y0 = (torch.rand((B, 1, d), device="cuda", generator=gen) * 2 - 1)
    t = np.linspace(0, 1, 17).astype(np.float32)
    dW = torch.randn((16, B, d), device="cuda", generator=gen) * 0.25
    a = px.sdeint(f, g, y0, t, px.Euler, options={"bm_increments": dW, "math": "tensor"})
    b = px.sdeint(f, g, y0, t, px.Euler, options={"bm_increments": dW, "math": "fp32"})
    r = {"case": f"sde EM 2x({d}-{h}-{d}) B={B}", **err(a, b)}
    if time:
        r["ms_tensor"] = timeit(lambda: px.sdeint(f, g, y0, t, px.Euler, options={"bm_increments": dW, "math": "tensor", "out_stride": 16, "check_status": False}))
        r["ms_fp32"] = timeit(lambda: px.sdeint(f, g, y0, t, px.Euler, options={"bm_increments": dW, "math": "fp32", "out_stride": 16}))
    return r


if __name__ == "__main__":
    quick = "quick" in sys.argv
    if "bench" in sys.argv:
        print(json.dumps(ode(64, 256, 512, "RK4", "id")), flush=True)
        print(json.dumps(ode(64, 256, 1 << 17, "RK4", "id", steps=100, time=True)), flush=True)
        print(json.dumps(sde(32, 64, 1 << 21, time=True)), flush=True)
        sys.exit(0)
    print(json.dumps(ode(64, 256, 128, "Euler", "id", steps=1)), flush=True)
    print(json.dumps(ode(64, 256, 4096, "Euler", "id", steps=1)), flush=True)
    print(json.dumps(ode(64, 256, 100, "RK4", "id")), flush=True)
    if not quick:
        for d, h in [(64, 128), (64, 64), (32, 256), (32, 128), (32, 64), (16, 64)]:
            print(json.dumps(ode(d, h, 333, "RK4", "cube" if d < 64 else "id")), flush=True)
            print(json.dumps(ode(d, h, 129, "Euler", "square")), flush=True)
        for d, h in [(32, 64), (32, 128), (64, 64), (16, 64)]:
            print(json.dumps(sde(d, h, 1000)), flush=True)
        print(json.dumps(ode(64, 256, 1 << 17, "RK4", "id", steps=100, time=True)), flush=True)
        print(json.dumps(sde(32, 64, 1 << 21, time=True)), flush=True)
