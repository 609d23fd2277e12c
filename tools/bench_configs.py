"""Secondary measurements (not the bench.py line): cfg3 / cfg4 / cfg5 kernels, device-resident, CUDA events.
Prints one JSON object per config with throughput and the roofline fractions of SURVEY 8(d)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import paddlexde_b200 as px
from paddlexde_b200.xde.base_dde import history_gather, history_gather_bwd
from tests.problems import fanin_weights

_PEAKS = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
_P = json.load(open(_PEAKS)) if os.path.exists(_PEAKS) else {}
HBM = _P.get("hbm_gbs", 6650.0)
TENSOR = _P.get("bf16_tflops_sustained", 1371.6)  # dense 16-bit tensor TFLOP/s (the tensor path runs kind::f16)
FFMA = 72.3  # TFLOP/s: tools/probe_fp32.py on this pool; bench.py overwrites it with the probe it timed in the same run


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def cfg3(B=1 << 17, math="tensor"):
    d, h = 64, 256
    field = px.MLPField(*fanin_weights(d, h, seed=1), pre="id")
    y0 = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, (B, 1, d)).astype(np.float32)).cuda()
    t = np.linspace(0, 1, 101).astype(np.float32)
    xde = px.xde.BaseODE(field, y0, t)
    s = px.RK4(xde=xde, y0=y0, rtol=1e-7, atol=1e-9, out_stride=10, math=math, check_status=False)  # kernel timing: stay asynchronous
    ms = timeit(lambda: s.integrate(t))
    steps = B * 100
    flops = steps * 16 * d * h
    byts = steps * 8 * d + B * 11 * d * 4
    r = {"config": "cfg3 rk4 64-256-64", "math": math, "B": B, "ms": ms, "traj_steps_per_s": steps / ms * 1e3,
         "tflops_algorithmic": flops / ms / 1e9, "hbm_gbs_algorithmic": byts / ms / 1e6,
         "frac_hbm": byts / ms / 1e6 / HBM}
    if math == "tensor":  # three fp16 MMAs per algorithmic product (hi*hi + hi*lo + lo*hi)
        r.update({"frac_tensor_peak_algorithmic": flops / ms / 1e9 / TENSOR,
                  "frac_tensor_peak_issued": 3 * flops / ms / 1e9 / TENSOR, "tensor_peak_tflops": TENSOR,
                  "bound": "FP32/ALU issue of the tanh epilogue (1 tanh per 256 algorithmic FLOP), see DESIGN 5.4b"})
    else:
        r["frac_ffma_peak"] = flops / ms / 1e9 / FFMA
    return r


def cfg4(B=1 << 21, math="tensor", generated=False):
    d, h = 32, 64
    f = px.MLPField(*fanin_weights(d, h, seed=2), pre="cube")
    g = px.MLPField(*fanin_weights(d, h, seed=3), pre="square")
    gen = torch.Generator(device="cuda").manual_seed(2)
    y0 = (torch.rand((B, 1, d), device="cuda", generator=gen) * 2 - 1)
    t = np.linspace(0, 1, 17).astype(np.float32)
    dW = None if generated else torch.randn((16, B, d), device="cuda", generator=gen) * 0.25
    xde = px.xde.BaseSDE(f, g, y0, t, bm_seed=2) if generated else px.xde.BaseSDE(f, g, y0, t, bm_increments=dW)
    s = px.Euler(xde=xde, y0=y0, rtol=1e-7, atol=1e-9, out_stride=16, math=math, check_status=False)
    ms = timeit(lambda: s.integrate(t))
    steps = B * 16
    flops = steps * (4 * d * h * 2)
    byts = steps * 12 * d  # read y, read dW, write y per trajectory-step
    # what a FUSED solve has to move: the increments (0 when generated in the kernel) + y0 in + the 2 emitted rows out
    moved = (0 if generated else steps * 4 * d) + B * d * 4 + B * 2 * d * 4
    return {"config": "cfg4 sde-EM 2x(32-64-32)" + (" increments generated in-kernel (Philox)" if generated else
                                                     " supplied increments"),
            "math": math, "B": B, "ms": ms,
            "traj_steps_per_s": steps / ms * 1e3, "tflops_algorithmic": flops / ms / 1e9,
            "hbm_gbs_required": moved / ms / 1e6, "frac_hbm": moved / ms / 1e6 / HBM,
            "hbm_gbs_survey_formula": byts / ms / 1e6, "frac_hbm_survey_formula": byts / ms / 1e6 / HBM,
            "note": "frac_hbm counts the bytes a fused solve must move (dW table + y0 + emitted rows); the SURVEY 8(d) "
                    "formula (12*D B per trajectory-step) also charges a y read+write per step that the fused kernel "
                    "never performs (r1 quoted that figure: inflated 2.6x).  B=2^22 is BASELINE's 4M (dW table 8 GiB)"
                    if B == 1 << 22 else "reduced batch (BASELINE's cfg4 is B=2^22)"}


def cfg5(Bh=4096, kind="cubic"):
    rng = np.random.default_rng(5)
    his = torch.from_numpy(rng.uniform(-1, 1, (Bh, 307, 288, 3)).astype(np.float32)).cuda()
    span = torch.arange(288, dtype=torch.float32, device="cuda")
    lags = torch.from_numpy((np.arange(12) + rng.uniform(0, 1, 12)).astype(np.float32)).cuda()
    # 20 calls between the events: a 60 us kernel is shorter than the host side of one call (two allocations +
    # ctypes), back-to-back launches queue up and the GPU time per call is what the events measure
    NB = 20
    ms = timeit(lambda: [history_gather(lags, his, span, kind) for _ in range(NB)]) / NB
    ms_single = timeit(lambda: history_gather(lags, his, span, kind))
    n = Bh * 307 * 12 * 3
    v, dv = history_gather(lags, his, span, kind)
    gy = torch.randn_like(v)
    ms_b = timeit(lambda: [history_gather_bwd(gy, dv) for _ in range(NB)]) / NB
    return {"config": f"cfg5 history gather {kind}", "B": Bh, "ms_fwd": ms, "ms_bwd": ms_b, "elements": n,
            "ms_fwd_single_call_incl_host": ms_single,
            "hbm_gbs_algorithmic_fwd": n * 20 / ms / 1e6, "frac_hbm_fwd": n * 20 / ms / 1e6 / HBM,
            "hbm_gbs_algorithmic_bwd": n * 8 / ms_b / 1e6, "frac_hbm_bwd": n * 8 / ms_b / 1e6 / HBM,
            "note": "his is 4.2 GB at B=4096 (the kernel must outlast the ~75 us host side of a call to be timed from the "
                    "host) but only 13 of 288 time rows are touched: 20 B/element algorithmic"}


def cfg5_real_size(kind="cubic"):
    """cfg5 at the size D3STN really calls it with (example/D3STN/train_dde.py: batch 8, 307 graph nodes, 288 history
    steps, 3 channels, 12 lags): 88 K output elements -- launch-bound.  Latency of one call through the shim (host
    side included) and of the kernel alone (20 queued launches between the events)."""
    rng = np.random.default_rng(5)
    his = torch.from_numpy(rng.uniform(-1, 1, (8, 307, 288, 3)).astype(np.float32)).cuda()
    span = torch.arange(288, dtype=torch.float32, device="cuda")
    lags = torch.from_numpy((np.arange(12) + rng.uniform(0, 1, 12)).astype(np.float32)).cuda()
    NB = 20
    ms = timeit(lambda: [history_gather(lags, his, span, kind) for _ in range(NB)]) / NB
    t0 = time.perf_counter()
    for _ in range(200):
        history_gather(lags, his, span, kind)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 200
    n = 8 * 307 * 12 * 3
    return {"config": f"cfg5 history gather {kind}, D3STN call size (8 x 307 rows, 288 steps, 12 lags, 3 channels)",
            "elements": n, "us_per_call_gpu_queued": ms * 1e3, "us_per_call_host_wall": wall * 1e6,
            "note": "88 K elements = 1.8 MB algorithmic: a few microseconds of kernel; the call is bound by launch + host "
                    "overhead (two allocations + one ctypes call), not by HBM"}


def cfg1_latency():
    """cfg1 = BASELINE configs[0]: example/ode_demo.py's training call, odeint(func, batch_y0 [20, 2], batch_t [10],
    Dopri5) with the 2-50-2 field at rtol 1e-7 -- 20 trajectories: pure latency.  One call through the public API from
    device-resident inputs to a synchronised result (host wall clock), and the kernel alone (20 queued solves between
    two events); with the reference's global controller and with one controller per trajectory."""
    from tests.problems import cfg2_tspan, cfg2_y0, spiral_weights
    field = px.MLPField(*spiral_weights(), pre="cube")
    y0 = torch.from_numpy(cfg2_y0(20)).cuda()
    t = cfg2_tspan(10)
    out = {"config": "cfg1 ode_demo call: dopri5 rtol 1e-7, MLP 2-50-2, batch 20, 10 output times (latency)"}
    for ctrl in ("batch", "trajectory"):
        call = lambda: px.odeint(field, y0, t, px.Dopri5, options={"controller": ctrl, "check_status": False})  # noqa: E731
        NB = 20
        ms = timeit(lambda: [call() for _ in range(NB)]) / NB
        t0 = time.perf_counter()
        for _ in range(100):
            call()
            torch.cuda.synchronize()
        out[f"us_per_call_gpu_queued_{ctrl}"] = ms * 1e3
        out[f"us_per_call_host_wall_synchronised_{ctrl}"] = (time.perf_counter() - t0) / 100 * 1e6
    return out


def cfg2_batch(B=1 << 20, norm="mixed"):
    """cfg2 with the REFERENCE-FAITHFUL controller: one global RMS norm and dt for the whole batch
    (utils/ode_utils.py:8-9), adjoint with the reference's default mixed norm or the seminorm."""
    from tests.problems import cfg2_tspan, cfg2_y0, spiral_weights
    tw = [torch.tensor(a, device="cuda", requires_grad=True) for a in spiral_weights()]
    field = px.MLPField(*tw, pre="cube")
    y0 = torch.from_numpy(cfg2_y0(B)).cuda()
    t = torch.from_numpy(cfg2_tspan(10))
    opts = {"controller": "batch"}
    aopts = {"controller": "batch"} if norm == "mixed" else {"controller": "batch", "norm": "seminorm"}
    st = {}

    def fwd():
        st["sol"] = px.odeint_adjoint(field, y0, t, solver=px.Dopri5, options=opts, adjoint_options=aopts)

    def bwd():
        for p in tw:
            p.grad = None
        fwd()
        st["sol"][-1].abs().mean().backward()

    ms_f = timeit(fwd, warm=4)
    ms = timeit(bwd, warm=6)  # the first cooperative launches and autograd passes are several times slower
    h = px.odeint_adjoint.last
    att_f = int(h["fwd_solver"].read_stats().n_attempts)
    att_b = int(h["bwd_stats"].read().n_attempts)
    return {"config": f"cfg2 dopri5 fwd+adjoint, controller=batch (reference-faithful global dt), adjoint norm {norm}",
            "B": B, "ms_fwd": ms_f, "ms_fwd_plus_adjoint": ms, "trajectory_steps": att_f + att_b,
            "traj_steps_per_s": (att_f + att_b) / ms * 1e3,
            "note": "one cooperative launch per solve, ~2 grid.sync per attempt; bit-exact dt / ratio / accept sequence vs the "
                    "oracle's batch run (forward, seminorm adjoint)"}


def cfg3_dopri5(B=1 << 15):
    """The cfg3 field (MLP 64-256-64) under the ADAPTIVE solver: dopri5 rtol 1e-6, one controller per trajectory,
    register-tiled FP32 field (xde_tile_adaptive.cu; bit-exact vs the oracle)."""
    d, h = 64, 256
    field = px.MLPField(*fanin_weights(d, h, seed=1), pre="id")
    y0 = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, (B, d)).astype(np.float32)).cuda()
    t = np.linspace(0, 1, 11).astype(np.float32)
    xde = px.xde.BaseODE(field, y0, t)
    s = px.Dopri5(xde=xde, y0=y0, rtol=1e-6, atol=1e-8, check_status=False)
    ms = timeit(lambda: s.integrate(t))
    st = s.read_stats()
    steps = int(st.n_attempts)
    flops = int(st.nfe) * 4 * d * h
    return {"config": "cfg3 field 64-256-64, dopri5 rtol=1e-6 (adaptive, per-trajectory controller)", "math": "fp32",
            "B": B, "ms": ms, "trajectory_steps": steps, "traj_steps_per_s": steps / ms * 1e3,
            "attempts_per_trajectory": steps / B, "tflops_algorithmic": flops / ms / 1e9,
            "frac_ffma_peak": flops / ms / 1e9 / FFMA,
            "note": "tiles of 32 trajectories advance together until the slowest row is done (no refill inside a tile)"}


def cfg3_adjoint(B=1 << 14):
    """The cfg3 field (MLP 64-256-64) trained with odeint_adjoint: dopri5 forward + adjoint, rtol 1e-6, one controller per
    trajectory (forward: xde_tile_adaptive.cu, backward: xde_adj_tile.cu -- FP32 tiles, gradients in tensor memory)."""
    from paddlexde_b200.functional.odeint_adjoint import adjoint_backward
    d, h = 64, 256
    field = px.MLPField(*fanin_weights(d, h, seed=1), pre="id")
    y0 = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, (B, d)).astype(np.float32)).cuda()
    t = np.linspace(0, 1, 5).astype(np.float32)
    xde = px.xde.BaseODE(field, y0, t)
    s = px.Dopri5(xde=xde, y0=y0, rtol=1e-6, atol=1e-8, controller="trajectory", check_status=False)
    sol = s.integrate(t)
    gy = torch.zeros_like(sol)
    gy[-1] = torch.sign(sol[-1]) / sol[-1].numel()
    ms_f = timeit(lambda: s.integrate(t))
    st = {}

    def bwd():
        st["r"] = adjoint_backward(field, t, sol, gy, rtol=1e-6, atol=1e-8, check_status=False)

    ms_b = timeit(bwd)
    a = st["r"][2].read()
    f = s.read_stats()
    evals = 6 * a.n_attempts + 3 * (t.size - 1) * B  # f0 + probe + the theta-only pass of f0 per segment
    flops = evals * 12 * d * h  # 4 state GEMMs + 2 gradient GEMMs of 2*D*H each
    return {"config": "cfg3 field 64-256-64, dopri5 rtol=1e-6 forward + adjoint (per-trajectory controller, seminorm)",
            "math": "fp32", "B": B, "ms_fwd": ms_f, "ms_adjoint": ms_b, "trajectory_steps_fwd": int(f.n_attempts),
            "trajectory_steps_adjoint": int(a.n_attempts), "status": [int(f.status), int(a.status)],
            "traj_steps_per_s": (f.n_attempts + a.n_attempts) / (ms_f + ms_b) * 1e3,
            "tflops_adjoint": flops / ms_b / 1e9, "frac_ffma_peak_adjoint": flops / ms_b / 1e9 / FFMA}


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg3", "cfg4", "cfg5"]
    for w in which:
        if w in ("cfg3", "cfg4"):
            for math in ("tensor", "fp32"):
                print(json.dumps(globals()[w](math=math)), flush=True)
            if w == "cfg4":
                print(json.dumps(cfg4(math="tensor", generated=True)), flush=True)
                print(json.dumps(cfg4(B=1 << 22, math="tensor", generated=True)), flush=True)
        elif w == "cfg3_dopri5":
            print(json.dumps(cfg3_dopri5()), flush=True)
        elif w == "cfg3_adjoint":
            print(json.dumps(cfg3_adjoint()), flush=True)
        elif w == "cfg2_batch":
            for norm in ("mixed", "seminorm"):
                print(json.dumps(cfg2_batch(norm=norm)), flush=True)
        else:
            print(json.dumps(globals()[w]()), flush=True)
