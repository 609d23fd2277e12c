"""The reference's example/ode_demo.py on the B200 path: fit the 2-50-2 neural ODE `dy/dt = net(y**3)` to the spiral
`dy/dt = y**3 A` from mini-batches of 20 trajectories x 10 output times (BASELINE config 1; with --batch 1048576 it is the
shape of config 2).  Same flow as the reference's script -- `odeint_adjoint(func, batch_y0, t_span, solver=Dopri5)`,
`loss = mean|pred - true|`, `loss.backward()`, RMSprop -- with the two solves running as fused sm_100a kernels and the
parameter gradients coming from the augmented reverse-time solve.

    python examples/ode_demo.py [--steps 200] [--batch 20] [--pred-len 10] [--controller trajectory|batch]

PyTorch plays the role Paddle plays in the reference (parameters, autograd tape, optimizer); INTEGRATION.md has the
Paddle-side binding."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import paddlexde_b200 as px  # noqa: E402


def spiral(n=1000):
    """SimpleDemoData (example/demo_utils.py:147-164): the true trajectory on linspace(0, 25, n), RK4 in float64."""
    A = np.array([[-0.1, 2.0], [-2.0, -0.1]])
    t = np.linspace(0.0, 25.0, n)
    y = np.empty((n, 2))
    y[0] = [2.0, 0.0]
    f = lambda v: (v ** 3) @ A  # noqa: E731
    for i in range(n - 1):
        h = t[i + 1] - t[i]
        k1 = f(y[i]); k2 = f(y[i] + 0.5 * h * k1); k3 = f(y[i] + 0.5 * h * k2); k4 = f(y[i] + h * k3)  # noqa: E702
        y[i + 1] = y[i] + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    return t.astype(np.float32), y.astype(np.float32)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--batch", type=int, default=20)
    ap.add_argument("--pred-len", type=int, default=10)
    ap.add_argument("--controller", default="trajectory", choices=["trajectory", "batch"])
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--seed", type=int, default=42)
    args = ap.parse_args(argv)
    rng = np.random.default_rng(args.seed)
    t_all, y_all = spiral()
    dev = torch.device("cuda")
    # ODEFunc (example/ode_demo.py:17-33): Linear(2, 50) -> Tanh -> Linear(50, 2) on y**3, W ~ 0.1 N(0, 1), b = 0;
    # Paddle's [in, out] weight layout
    params = [torch.tensor(a, device=dev, requires_grad=True) for a in
              ((0.1 * rng.standard_normal((2, 50))).astype(np.float32), np.zeros(50, np.float32),
               (0.1 * rng.standard_normal((50, 2))).astype(np.float32), np.zeros(2, np.float32))]
    func = px.MLPField(*params, pre="cube")
    opt = torch.optim.RMSprop(params, lr=args.lr)
    t_span = t_all[:args.pred_len]  # the reference integrates every mini-batch over the first pred_len grid times
    losses = []
    for step in range(1, args.steps + 1):
        s = rng.integers(0, len(t_all) - args.pred_len, args.batch)
        batch_y0 = torch.from_numpy(y_all[s]).to(dev)                                            # [B, D]
        batch_y = torch.from_numpy(np.stack([y_all[s + i] for i in range(args.pred_len)])).to(dev)  # [T, B, D]
        pred_y = px.odeint_adjoint(func, batch_y0, t_span, solver=px.Dopri5, options={"controller": args.controller})
        loss = (pred_y - batch_y).abs().mean()
        opt.zero_grad()
        loss.backward()
        opt.step()  # in-place update: the field re-reads its device copies on the next solve (version counters)
        losses.append(float(loss))
        if step % max(args.steps // 10, 1) == 0:
            print(f"Iter {step:04d} | Total Loss {losses[-1]:.6f}")
    return losses, params


if __name__ == "__main__":
    main()
