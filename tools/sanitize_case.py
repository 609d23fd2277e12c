"""Small all-kernel exercise for compute-sanitizer (memcheck / racecheck): every C-ABI entry once."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import paddlexde_b200 as px
from paddlexde_b200.functional.odeint_adjoint import adjoint_backward
from paddlexde_b200.xde.base_dde import history_gather, history_gather_bwd
from tests.problems import cfg2_tspan, cfg2_y0, fanin_weights, spiral_weights

f = px.MLPField(*spiral_weights(), pre="cube")
y0 = torch.from_numpy(cfg2_y0(300)).cuda()
t = cfg2_tspan(5)
sol = px.odeint(f, y0, t, px.Dopri5)
gy = torch.zeros_like(sol); gy[-1] = torch.sign(sol[-1]) / sol[-1].numel()
g, a0, st, _ = adjoint_backward(f, t, sol, gy, return_adj_y0=True)
xde = px.xde.BaseODE(f, y0, t)
px.Dopri5(xde=xde, y0=y0, rtol=1e-7, atol=1e-9, controller="batch").integrate(t)
px.odeint(f, y0.reshape(300, 1, 2), t, px.RK4); px.odeint(f, y0.reshape(300, 1, 2), t, px.Euler)
w3 = [3.0 * a for a in fanin_weights(2, 50, seed=5)]
f3 = px.MLPField(*w3, pre="id")
s3 = px.odeint(f3, y0 * 0.3, np.linspace(0, 4, 5, dtype=np.float32), px.Dopri5, rtol=1e-6, atol=1e-8)
adjoint_backward(f3, np.linspace(0, 4, 5, dtype=np.float32), s3, 0.01 * torch.randn_like(s3), rtol=1e-6, atol=1e-8)
fd, gd = px.MLPField(*fanin_weights(4, 32, seed=2), pre="cube"), px.MLPField(*fanin_weights(4, 32, seed=3), pre="square")
dW = 0.25 * torch.randn(16, 200, 4, device="cuda")
for sch in ("em", "milstein"):
    px.sdeint(fd, gd, torch.rand(200, 1, 4, device="cuda"), np.linspace(0, 1, 17, dtype=np.float32), px.Euler,
              options={"bm_increments": dW, "scheme": sch})
ft = px.MLPField(*fanin_weights(64, 256, seed=1), pre="id")
px.odeint(ft, torch.rand(70, 1, 64, device="cuda"), np.linspace(0, 1, 4, dtype=np.float32), px.RK4)
f32_, g32 = px.MLPField(*fanin_weights(32, 64, seed=2), pre="cube"), px.MLPField(*fanin_weights(32, 64, seed=3), pre="square")
px.sdeint(f32_, g32, torch.rand(100, 1, 32, device="cuda"), np.linspace(0, 1, 5, dtype=np.float32), px.Euler,
          options={"bm_increments": 0.25 * torch.randn(4, 100, 32, device="cuda")})
his = torch.rand(4, 37, 288, 3, device="cuda"); span = torch.arange(288.0, device="cuda")
lags = torch.arange(12.0, device="cuda") + 0.3
for kind in ("linear", "cubic"):
    v, d = history_gather(lags, his, span, kind)
    history_gather_bwd(torch.randn_like(v), d)
torch.cuda.synchronize()
print("sanitize case done; launches:", px.launch_count())
