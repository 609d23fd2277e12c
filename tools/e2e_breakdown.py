"""Diagnostic: where the end-to-end (public API, host buffers) step spends its time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import paddlexde_b200 as px
from bench import workload

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
w, y0_np, t = workload(B)
dev = torch.device("cuda", 0)
tw = [torch.tensor(a, device=dev, requires_grad=True) for a in w]
field = px.MLPField(*tw, pre="cube")
y0_pin = torch.from_numpy(y0_np).pin_memory()
y0_buf = torch.empty((B, 2), device=dev)
th = torch.from_numpy(t)

def tick(label, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"  {label:28s} {1e3*(t1-t0):8.2f} ms")
    return t1

for it in range(4):
    print("iter", it)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for p in tw: p.grad = None
    y0_buf.copy_(y0_pin, non_blocking=True)
    t0 = tick("h2d", t0)
    sol = px.odeint_adjoint(field, y0_buf, th, solver=px.Dopri5, options={"controller": "trajectory"})
    t0 = tick("odeint_adjoint fwd", t0)
    loss = sol[-1].abs().mean()
    t0 = tick("loss", t0)
    loss.backward()
    t0 = tick("backward", t0)
    flat = torch.cat([p.grad.reshape(-1) for p in tw]).cpu(); l = loss.item()
    t0 = tick("d2h", t0)
