"""One launch set of the history gather at the scaled cfg5 size (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from paddlexde_b200.xde.base_dde import history_gather, history_gather_bwd
rng = np.random.default_rng(5)
his = torch.from_numpy(rng.uniform(-1, 1, (4096, 307, 288, 3)).astype(np.float32)).cuda()
span = torch.arange(288, dtype=torch.float32, device="cuda")
lags = torch.from_numpy((np.arange(12) + rng.uniform(0, 1, 12)).astype(np.float32)).cuda()
for _ in range(3):
    v, dv = history_gather(lags, his, span, sys.argv[1] if len(sys.argv) > 1 else "cubic")
    g = history_gather_bwd(torch.ones_like(v), dv)
torch.cuda.synchronize()
