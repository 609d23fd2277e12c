"""One launch of the tensor-core kernels at cfg3 / cfg4 size (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import paddlexde_b200 as px
from tests.problems import fanin_weights

which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
math = sys.argv[2] if len(sys.argv) > 2 else "tensor"
if which == "cfg3":
    d, h, B = 64, 256, 1 << 17
    field = px.MLPField(*fanin_weights(d, h, seed=1), pre="id")
    y0 = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, (B, 1, d)).astype(np.float32)).cuda()
    t = np.linspace(0, 1, 101).astype(np.float32)
    for _ in range(2):
        px.odeint(field, y0, t, px.RK4, options={"math": math, "out_stride": 10, "check_status": False})
else:
    d, h, B = 32, 64, 1 << 21
    f = px.MLPField(*fanin_weights(d, h, seed=2), pre="cube")
    g = px.MLPField(*fanin_weights(d, h, seed=3), pre="square")
    gen = torch.Generator(device="cuda").manual_seed(2)
    y0 = (torch.rand((B, 1, d), device="cuda", generator=gen) * 2 - 1)
    t = np.linspace(0, 1, 17).astype(np.float32)
    dW = torch.randn((16, B, d), device="cuda", generator=gen) * 0.25
    for _ in range(2):
        px.sdeint(f, g, y0, t, px.Euler, options={"bm_increments": dW, "math": math, "out_stride": 16, "check_status": False})
torch.cuda.synchronize()
