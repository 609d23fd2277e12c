"""bench.py --config cfg3|cfg4|cfg5: the other BASELINE.json configs with the bench line schema of cfg2.

  cfg3  batched neural ODE ensemble, 64-D state, MLP 64-256-64, RK4 (3/8 rule) on a fixed grid of 100 steps,
        2^17 trajectories per GPU (weak; 8 GPUs = the named 2^20) or 2^20 in total (--scaling strong);
        tcgen05 field (csrc/xde_tc.cu).  metric: rk4 trajectory-steps/s; roofline: tensor.
  cfg4  Euler-Maruyama sdeint, 32-D state, drift/diffusion MLP 32-64-32, supplied Brownian increments, batch 4M
        (2^22, the dW table is 8 GiB in HBM), 16 steps.  metric: sde trajectory-steps/s; roofline: hbm on the
        bytes a fused solve must move.
  cfg5  D3STN history gather (cubic Hermite, 307 nodes x 288 steps x 3 channels, 12 learnable lags): bandwidth at
        a scaled batch of 4096 graphs + latency at the real call size (batch 8).  metric: gathered elements/s.

value = device-resident inputs, CUDA events on the launching stream, a 256 MiB L2 flush between steps (inside the
timed region), max over ranks.  e2e = the same through the public entry point from pinned HOST buffers (H2D of
the step's inputs and D2H of its result summary inside the timed region).  Trajectories shard by batch; there
is no collective on these paths."""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from tests.problems import fanin_weights  # noqa: E402

UNIT = "trajectory-steps/s"


def spec(args, world):
    c = args.config
    if c == "cfg3":
        Bg = (1 << 20) if args.scaling == "strong" else (1 << 17) * world
        return dict(name="cfg3", metric="rk4 trajectory-steps/s (fixed grid, 64-256-64)", d=64, h=256, n_steps=100,
                    global_batch=Bg, workload="cfg3: odeint RK4 (3/8 rule) over linspace(0,1,101), MLP 64-256-64 (id), "
                                               "every 10th grid point emitted, tcgen05 field (fp16-split 3-product GEMMs)")
    if c == "cfg4":
        Bg = (1 << 22) if args.scaling == "strong" else (1 << 22) * world
        return dict(name="cfg4", metric="sde-EM trajectory-steps/s (supplied increments)", d=32, h=64, n_steps=16,
                    global_batch=Bg, workload="cfg4: sdeint Euler-Maruyama over linspace(0,1,17), drift 32-64-32 (y**3) and "
                                               "diffusion 32-64-32 (y**2), supplied increments dW [16, B, 32], last row emitted")
    return dict(name="cfg5", metric="history-gather elements/s (cubic Hermite)", d=3, h=0, n_steps=1,
                global_batch=4096 * world, workload="cfg5: HistoryIndex (CubicHermiteSpline evaluate + derivative), his "
                                                    "[B, 307, 288, 3], 12 lags; B = 4096 graphs (scaled) and B = 8 (D3STN)")


def reference_arm(args):
    """The reference's algorithm for the config on the host cores (C oracle port; rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import xde_oracle as xo

    xo.build()
    sp = spec(args, 1)
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(1)
    if sp["name"] == "cfg3":
        Bs = 1 << 10
        om = xo.MLP(*fanin_weights(64, 256, seed=1), "id")
        y0 = rng.uniform(-1, 1, (Bs, 64)).astype(np.float32)
        t = np.linspace(0, 1, 101).astype(np.float32)
        fn = lambda: xo.fixed_mlp("rk4", om, y0, t, nthreads=cores)
        units = Bs * 100
    elif sp["name"] == "cfg4":
        Bs = 1 << 14
        f, g = xo.MLP(*fanin_weights(32, 64, seed=2), "cube"), xo.MLP(*fanin_weights(32, 64, seed=3), "square")
        y0 = rng.uniform(-1, 1, (Bs, 32)).astype(np.float32)
        t = np.linspace(0, 1, 17).astype(np.float32)
        dW = (0.25 * rng.standard_normal((16, Bs, 32))).astype(np.float32)
        fn = lambda: xo.sde_mlp("em", f, g, y0, t, dW, nthreads=cores)
        units = Bs * 16
    else:
        Bs = 64
        his = rng.uniform(-1, 1, (Bs, 307, 288, 3)).astype(np.float32)
        span = np.arange(288, dtype=np.float32)
        lags = (np.arange(12) + rng.uniform(0, 1, 12)).astype(np.float32)
        fn = lambda: xo.history_gather("cubic", his, span, lags)
        units = Bs * 307 * 12 * 3
        cores = 1
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    v = units * args.steps / dt
    unit = UNIT if sp["name"] != "cfg5" else "elements/s"
    sample = f"each step = {units} units of {sp['name']} (B = {Bs}) on {cores} host threads, C oracle port of the reference algorithm"
    print(json.dumps({"impl": "reference", "metric": sp["metric"], "value": v, "unit": unit, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
                      "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": {"workload": sp["workload"], "sample_batch": Bs},
                      "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


def main(args, ClockSampler, peaks):
    if args.impl == "reference":
        return reference_arm(args)
    import torch
    import torch.distributed as dist

    import paddlexde_b200 as px
    from paddlexde_b200 import distributed as pxd
    from paddlexde_b200.xde.base_dde import history_gather

    rank, world, local = pxd.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sp = spec(args, world)
    lo, hi = pxd.shard_rows(sp["global_batch"], rank, world)
    B = hi - lo
    d, h = sp["d"], sp["h"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gen = torch.Generator(device="cpu").manual_seed(100 + lo % 9973)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    extra = {}

    if sp["name"] == "cfg3":
        field = px.MLPField(*fanin_weights(d, h, seed=1), pre="id")
        y0_pin = (torch.rand((B, 1, d), generator=gen) * 2 - 1).pin_memory()
        y0 = y0_pin.to(dev)
        t = np.linspace(0, 1, 101).astype(np.float32)
        mk = lambda y: px.RK4(xde=px.xde.BaseODE(field, y, t), y0=y, rtol=1e-7, atol=1e-9, out_stride=10, math="tensor",
                              check_status=False)
        dev_step = lambda: mk(y0).integrate(t)
        h2d_bytes, d2h_bytes = y0_pin.numel() * 4, 4

        def e2e_step():
            y = y0_pin.to(dev, non_blocking=True)
            return float(px.odeint(field, y, t, px.RK4, options={"out_stride": 10})[:, -1].abs().mean().item())
        units = B * 100
        flops, req_bytes = units * 16 * d * h, B * d * 4 + B * 11 * d * 4
        kernel = "tc::fixed_tc_kernel<64,256,RK4>"
    elif sp["name"] == "cfg4":
        f = px.MLPField(*fanin_weights(d, h, seed=2), pre="cube")
        g = px.MLPField(*fanin_weights(d, h, seed=3), pre="square")
        t = np.linspace(0, 1, 17).astype(np.float32)
        y0_pin = (torch.rand((B, 1, d), generator=gen) * 2 - 1).pin_memory()
        y0 = y0_pin.to(dev)
        gd = torch.Generator(device=dev).manual_seed(2 + rank)
        dW = torch.randn((16, B, d), device=dev, generator=gd) * 0.25
        mk = lambda y, w: px.Euler(xde=px.xde.BaseSDE(f, g, y, t, bm_increments=w), y0=y, rtol=1e-7, atol=1e-9,
                                   out_stride=16, math="tensor", check_status=False)
        dev_step = lambda: mk(y0, dW).integrate(t)
        # e2e with HOST increments is PCIe-bound (8 GiB per step at B = 2^22); a bounded host table keeps the run short:
        # the first 2^19 trajectories' increments cross the bus every step, the timed batch is those 2^19 trajectories
        Be = min(B, 1 << 19)
        dW_pin = torch.empty((16, Be, d), pin_memory=True)
        dW_pin.copy_(dW[:, :Be])
        h2d_bytes, d2h_bytes = Be * d * 4 + dW_pin.numel() * 4, 4
        extra["e2e_batch"] = Be

        def e2e_step():
            y = y0_pin[:Be].to(dev, non_blocking=True)
            w = dW_pin.to(dev, non_blocking=True)
            return float(px.sdeint(f, g, y, t, px.Euler, options={"bm_increments": w, "out_stride": 16})[:, -1].abs().mean().item())
        units = B * 16
        flops, req_bytes = units * 8 * d * h, units * 4 * d + B * d * 4 + B * 2 * d * 4
        kernel = "tc::fixed_tc2_kernel<32,64,EM>"
    else:
        rng = np.random.default_rng(5 + rank)
        his = torch.from_numpy(rng.uniform(-1, 1, (B, 307, 288, 3)).astype(np.float32)).to(dev)
        span = torch.arange(288, dtype=torch.float32, device=dev)
        lags = torch.from_numpy((np.arange(12) + np.random.default_rng(5).uniform(0, 1, 12)).astype(np.float32)).to(dev)
        dev_step = lambda: history_gather(lags, his, span, "cubic")
        his8_pin = his[:8].cpu().pin_memory()
        h2d_bytes, d2h_bytes = his8_pin.numel() * 4, 8 * 307 * 12 * 3 * 4 * 2
        extra["e2e_batch"] = 8

        def e2e_step():  # the D3STN call: batch 8 from host memory, both outputs back to the host
            v, dv = history_gather(lags, his8_pin.to(dev, non_blocking=True), span, "cubic")
            return v.cpu(), dv.cpu()
        units = B * 307 * 12 * 3
        flops, req_bytes = 0, units * 20
        kernel = "history_gather_kernel<HERMITE>"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n, flush_l2=True):
        evs = []
        barrier()
        e0, e1 = ev(), ev()
        w0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            if flush_l2:
                flush.zero_()
            a, b = ev(), ev()
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        e1.record()
        host_ms = 1e3 * (time.perf_counter() - w0) / n  # host time to ENQUEUE a step (diagnostic)
        barrier()
        extra.setdefault("host_enqueue_ms_per_step", []).append(round(host_ms, 3))
        return e0.elapsed_time(e1), float(np.mean([a.elapsed_time(b) for a, b in evs]))

    W = max(args.warmup, 3)
    for _ in range(W):
        dev_step()
    uuid = local
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        pass
    sampler = ClockSampler(uuid) if rank == 0 else None
    n0 = px.launch_count()
    ms, k_ms = timed(dev_step, args.steps)
    launches = px.launch_count() - n0
    clocks = sampler.stop() if sampler else None
    for _ in range(3):
        e2e_step()
    ms_e2e, _ = timed(e2e_step, args.steps)
    if world > 1:
        tt = torch.tensor([ms, ms_e2e, k_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e, k_ms = tt.tolist()
    if rank == 0:
        hbm, which = peaks()
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        tensor_peak = pk.get("bf16_tflops_sustained", 1400.0)
        g_units = sp["global_batch"] * (units // B)
        e2e_units = g_units if "e2e_batch" not in extra else extra["e2e_batch"] * (units // B) * world
        unit = UNIT if sp["name"] != "cfg5" else "elements/s"
        if sp["name"] == "cfg3":
            roof = {"bound": "tensor", "kernel": kernel, "achieved": flops / (k_ms * 1e-3) / 1e12, "peak": tensor_peak,
                    "unit": "TFLOP/s", "frac": flops / (k_ms * 1e-3) / 1e12 / tensor_peak, "traffic": None,
                    "peak_source": "bf16_tflops_sustained, " + which,
                    "note": "algorithmic FLOPs = 16*D*H per trajectory-step (SURVEY 8(d)); the kernel issues 3 fp16 MMAs per "
                            "algorithmic product (hi*hi + hi*lo + lo*hi) to reach fp32 accuracy"}
        else:
            roof = {"bound": "hbm", "kernel": kernel, "achieved": req_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm,
                    "unit": "GB/s", "frac": req_bytes / (k_ms * 1e-3) / 1e9 / hbm, "traffic": None, "peak_source": which,
                    "note": "bytes a fused solve must move per launch: increments + y0 + emitted rows (cfg4); 20 B per "
                            "gathered element (cfg5: 3 neighbours read through L1, value + derivative written)"}
        out = {"metric": sp["metric"], "value": g_units * args.steps / (ms * 1e-3), "unit": unit, "n_gpus": world,
               "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": args.scaling, "vs_baseline": None, "dtype": "f32 (fp16-split tcgen05 MMAs, fp32 accumulate)"
               if sp["name"] != "cfg5" else "f32", "data": "synthetic",
               "config": {"workload": sp["workload"], "batch_per_gpu": B, "global_batch": sp["global_batch"],
                          "parallelism": f"batch-sharded x{world}, no collective",
                          "l2": "256 MiB buffer written between steps (inside the timed region)", **extra},
               "kernel_ms": k_ms, "roofline": roof,
               "e2e": {"value": e2e_units * args.steps / (ms_e2e * 1e-3), "unit": unit, "h2d_bytes_per_step": int(h2d_bytes),
                       "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": ms_e2e / args.steps},
               "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(out), flush=True)
    if world > 1:  # no teardown (see bench.py): every collective is done, leave with exit code 0
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
