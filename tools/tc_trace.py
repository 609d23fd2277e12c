"""Phase trace of one CTA of the tensor-core kernel (debug build of the library with -DXDE_TC_TRACE).
Build:  make -C paddlexde_b200/csrc trace      ->  tools/_trace/libxde_trace.so
Run:    python tools/tc_trace.py cfg3|cfg4      (prints cycle deltas per phase for a few evaluations)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import paddlexde_b200._lib as L
L._SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_trace", "libxde_trace.so")
import paddlexde_b200 as px
from tests.problems import fanin_weights

which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
CAP = 4096
buf = torch.zeros((2, CAP), dtype=torch.int64, device="cuda")
h = L.lib()
h.xde_tc_trace_set.argtypes = [C.c_void_p]
assert h.xde_tc_trace_set(buf.data_ptr()) == 0
if which == "cfg3":
    d, hh, B = 64, 256, 148 * 128
    field = px.MLPField(*fanin_weights(d, hh, seed=1), pre="id")
    y0 = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, (B, 1, d)).astype(np.float32)).cuda()
    t = np.linspace(0, 1, 21).astype(np.float32)
    px.odeint(field, y0, t, px.RK4, options={"math": "tensor", "out_stride": 20})
else:
    d, hh, B = 32, 64, 148 * 128 * 8
    f = px.MLPField(*fanin_weights(d, hh, seed=2), pre="cube")
    g = px.MLPField(*fanin_weights(d, hh, seed=3), pre="square")
    y0 = torch.rand((B, 1, d), device="cuda") * 2 - 1
    t = np.linspace(0, 1, 17).astype(np.float32)
    dW = torch.randn((16, B, d), device="cuda") * 0.25
    px.sdeint(f, g, y0, t, px.Euler, options={"bm_increments": dW, "math": "tensor", "out_stride": 16})
torch.cuda.synchronize()
tr = buf.cpu().numpy().reshape(2, CAP // 2, 2)
names = {0: "eval start", 1: "u stored+arrived", 2: "z0 ready, ld issued", 20: "f_ready seen", 21: "F read, eval end",
         100: "MMA: u_ready seen", 130: "MMA: L2 issued + commit f"}
names.update({3: "u STTM issued", 4: "u wait::st done", 200: "round start", 201: "H(A) done", 202: "H(B) done",
              210: "F(A) read", 211: "F(B) read", 220: "update+U(A) done", 221: "update+U(B) done"})
for c in range(4):
    names[30 + c] = f"chunk {c} STTM issued"; names[40 + c] = f"chunk {c} wait::st done"
    names[10 + c] = f"chunk {c} tanh stored+arrived"; names[110 + c] = f"MMA: L1 chunk {c} issued+commit"
    names[120 + c] = f"MMA: h_ready[{c}] seen"
ev = [(int(tag), int(clk), r) for r in range(2) for tag, clk in tr[r] if clk]
ev.sort(key=lambda x: x[1])
# print evaluations 8..10 of compute warp 0 (steady state) merged with the MMA thread's events
starts = [clk for tag, clk, r in ev if tag in (0, 200)]
lo, hi = starts[8], starts[11]
prev = lo
for tag, clk, r in ev:
    if lo <= clk < hi:
        print(f"{clk - lo:8d} (+{clk - prev:5d})  {'MMA ' if r else 'warp0'}  {names.get(tag, tag)}")
        prev = clk
per = np.diff(starts[4:60])
print("cycles per evaluation (warp 0): median", int(np.median(per)), "min", int(per.min()), "max", int(per.max()))
