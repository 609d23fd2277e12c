"""Does ordering the batch by state locality (less controller divergence inside a warp) pay?  cfg2 forward+adjoint."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import paddlexde_b200 as px
from tests.problems import cfg2_tspan, cfg2_y0, spiral_weights

B = 1 << 20
dev = torch.device("cuda")
tw = [torch.tensor(a, device=dev, requires_grad=True) for a in spiral_weights()]
field = px.MLPField(*tw, pre="cube")
y0_np = cfg2_y0(B)
t = torch.from_numpy(cfg2_tspan(10))


def morton(y, bits=10):
    q = []
    for d in range(y.shape[1]):
        v = y[:, d]
        q.append(np.clip(((v - v.min()) / (v.max() - v.min() + 1e-12) * ((1 << bits) - 1)).astype(np.uint32), 0, (1 << bits) - 1))
    code = np.zeros(y.shape[0], np.uint64)
    for b in range(bits):
        for d in range(y.shape[1]):
            code |= ((q[d] >> b) & 1).astype(np.uint64) << np.uint64(b * y.shape[1] + d)
    return code


orders = {
    "unsorted": np.arange(B),
    "sort y0[:,0]": np.argsort(y0_np[:, 0], kind="stable"),
    "sort |y0|": np.argsort(np.linalg.norm(y0_np, axis=1), kind="stable"),
    "morton(y0)": np.argsort(morton(y0_np), kind="stable"),
}
for name, perm in orders.items():
    y0 = torch.from_numpy(np.ascontiguousarray(y0_np[perm])).to(dev)

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
    print(f"{name:14s} fwd+adjoint ms (median of 5): {np.median(ts):.3f}  grad checksum {float(g.double().sum()):.12f}")
