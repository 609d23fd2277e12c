"""sdeint + sdeint_adjoint at scale (small state, cfg2's field shapes): kernel timings, CUDA events."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import paddlexde_b200 as px
from paddlexde_b200.functional.sdeint_adjoint import sde_adjoint_backward
from tests.problems import fanin_weights

B, d, h, T = 1 << 20, 2, 50, 17
f = px.MLPField(*fanin_weights(d, h, seed=2), pre="cube")
g = px.MLPField(*fanin_weights(d, h, seed=3), pre="square")
y0 = (torch.rand((B, 1, d), device="cuda") * 2 - 1)
t = np.linspace(0, 1, T).astype(np.float32)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), r


ms_f, sol = timeit(lambda: px.sdeint(f, g, y0, t, px.Euler, options={"bm_seed": 9}))
gy = torch.zeros_like(sol)
gy[:, -1] = torch.sign(sol[:, -1]) / sol[:, -1].numel()
ms_b, (gf, gg, _) = timeit(lambda: sde_adjoint_backward(f, g, t, sol, gy, bm_seed=9))
steps = B * (T - 1)
print(json.dumps({"config": f"sdeint + sdeint_adjoint, 2x({d}-{h}-{d}), B=2^20, {T - 1} EM steps, increments generated in-kernel",
                  "ms_forward": ms_f, "ms_adjoint": ms_b, "traj_steps_per_s_forward": steps / ms_f * 1e3,
                  "traj_steps_per_s_adjoint": steps / ms_b * 1e3, "finite": bool(torch.isfinite(gf).all() and torch.isfinite(gg).all())}))
