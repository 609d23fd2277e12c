"""Time cfg2 forward+adjoint with an alternative build of the library (kernel tuning experiments).
usage: python tools/adj_variants.py path/to/libxde_variant.so"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import paddlexde_b200._lib as L
if len(sys.argv) > 1:
    L._SO = os.path.abspath(sys.argv[1])
import paddlexde_b200 as px
from tests.problems import cfg2_tspan, cfg2_y0, spiral_weights

B = 1 << 20
dev = torch.device("cuda")
tw = [torch.tensor(a, device=dev, requires_grad=True) for a in spiral_weights()]
field = px.MLPField(*tw, pre="cube")
y0 = torch.from_numpy(cfg2_y0(B)).to(dev)
t = torch.from_numpy(cfg2_tspan(10))


def step():
    for p in tw:
        p.grad = None
    sol = px.odeint_adjoint(field, y0, t, solver=px.Dopri5)
    sol[-1].abs().mean().backward()
    return torch.cat([p.grad.reshape(-1) for p in tw])


for _ in range(3):
    g = step()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g = step(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(os.path.basename(L._SO), "fwd+adjoint ms (median of 5):", round(float(np.median(ts)), 3), "grad checksum", float(g.double().sum()))
