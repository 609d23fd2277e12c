"""Measure the FP32 pipe ceilings of this GPU: scalar FFMA and packed FFMA2 (TFLOP/s)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paddlexde_b200 import _lib, _tensor as T
lib = _lib.lib()
sink = torch.zeros(1, device="cuda")
for name in ("xde_probe_ffma_f32", "xde_probe_ffma2_f32"):
    fn = getattr(lib, name)
    n = C.c_int64(0)
    for _ in range(2):
        _lib.check(fn(1 << 14, T.ptr(sink), C.byref(n), T.stream()))
    torch.cuda.synchronize()
    best = 0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _lib.check(fn(1 << 16, T.ptr(sink), C.byref(n), T.stream())); e1.record()
        torch.cuda.synchronize()
        best = max(best, n.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    print(f"{name}: {best:.1f} TFLOP/s")
