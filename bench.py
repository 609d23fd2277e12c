#!/usr/bin/env python
"""bench.py -- dopri5 trajectory-steps/s of the B200 path on BASELINE.json configs[1] (cfg2):

    odeint_adjoint, Dopri5, rtol 1e-7 / atol 1e-9, MLP 2-50-2 (y**3 pre-activation),
    B = 2^20 trajectories per GPU, t_span = linspace(0, 25, 1000)[:10], loss = mean|y_T|.

A "step" is one forward solve + one adjoint (backward) solve of the whole batch.  A
trajectory-step is one solver step ATTEMPT (accepted or rejected) of one trajectory;
the attempt counts are read from the device-side counters of the kernels.

  value  inputs resident in HBM, kernels only (+ the loss-gradient fill and, for N > 1, the NCCL
         all-reduce of the 252 parameter gradients); CUDA events, max over ranks.
  e2e    through the public API (paddlexde_b200.odeint_adjoint + .backward()) from pinned HOST
         buffers: H2D of y0 (double-buffered on a copy stream: the copy for the next step overlaps this
         step's solve), forward, loss, backward, D2H of loss + parameter gradients, every step.
  roofline       dominant kernel vs the measured HBM copy bandwidth (MEASURED_PEAKS.json), with the
                 FP32-pipe figures that actually bound this field (D=2, H=50) beside it.
  cpu_baseline   the CPU oracle (port of the reference algorithm; Paddle is not installable) on all
                 host cores, on a bounded sample of the same workload.

`--impl reference` times that CPU oracle as the reference arm (the reference is pure Python on
Paddle, which cannot be installed offline here; DESIGN.md "Reference arm").
Multi-GPU: launched by torchrun, one rank per GPU, trajectories sharded by batch.  `--scaling weak` (default:
2^20 trajectories PER GPU) or `--scaling strong` (the 1M-trajectory batch split N ways).
`--config cfg3|cfg4|cfg5` measures the other BASELINE.json configs with the same line schema (tools/bench_lines.py).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "dopri5 trajectory-steps/s (forward + adjoint)"
UNIT = "trajectory-steps/s"
BYTES_FWD, BYTES_ADJ = 32, 64          # algorithmic HBM bytes per trajectory-step (SURVEY 8(d), D = 2)
FLOPS_FWD, FLOPS_ADJ = 2400, 7200      # algorithmic FLOPs per trajectory-step (24DH, 72DH)


def workload(B, seed=0, n_t=10):
    rng = np.random.default_rng(42)
    w1 = (0.1 * rng.standard_normal((2, 50))).astype(np.float32)
    w2 = (0.1 * rng.standard_normal((50, 2))).astype(np.float32)
    w = (w1, np.zeros(50, np.float32), w2, np.zeros(2, np.float32))
    y0 = (np.array([2.0, 0.0]) + 0.5 * np.random.default_rng(seed).standard_normal((B, 2))).astype(np.float32)
    t = np.linspace(0.0, 25.0, 1000).astype(np.float32)[:n_t]
    return w, y0, t


def latest_profile(kernel_substr, ms_measured, tol=0.02):
    """ncu evidence for the dominant kernel WITHOUT pasted constants: the newest profiles/r*_ncu_full_*.md whose
    section for `kernel_substr` was captured on a launch as long as the one this run timed (CUDA events vs
    gpu__time_duration within `tol`).  -> dict(traffic bytes, fma_pipe_active, duration_ms, source) or None."""
    import glob
    import re

    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full_*.md"))):
        sec = None
        for block in open(path).read().split("### ")[1:]:
            if kernel_substr in block.splitlines()[0]:
                sec = block
                break
        if sec is None:
            continue
        vals = {}
        for m in re.finditer(r"\| ([a-z_0-9.]+) \| ([-0-9.e+]+) \| ([^|]*) \|", sec):
            vals[m.group(1)] = (float(m.group(2)), m.group(3).strip())
        if "gpu__time_duration.sum" not in vals:
            continue
        dur, unit = vals["gpu__time_duration.sum"]
        dur_ms = dur * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(unit, 1.0)

        def nbytes(k):
            v, u = vals.get(k, (0.0, "byte"))
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)

        rec = {"duration_ms_ncu": dur_ms, "traffic": int(nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum")),
               "fma_pipe_active": vals.get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", (None,))[0],
               "issue_ipc": vals.get("sm__inst_issued.avg.per_cycle_active", (None,))[0],
               "source": "ncu --set full, " + os.path.relpath(path, ROOT),
               "matches_this_run": abs(dur_ms - ms_measured) <= tol * ms_measured}
        if rec["matches_this_run"] or best is None or not best["matches_this_run"]:
            if rec["matches_this_run"] or best is None:
                best = rec
    return best


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------
# CPU oracle legs (cpu_baseline, --impl reference)
# ---------------------------------------------------------------------------------------------------
def oracle_pass(xo, om, y0, t, nthreads):
    """One forward + adjoint pass of the CPU oracle; returns (attempts, seconds)."""
    t0 = time.perf_counter()
    sol, st, _, rc = xo.dopri5_mlp(om, y0, t, nthreads=nthreads)
    gy = np.zeros_like(sol)
    gy[-1] = np.sign(sol[-1]) / sol[-1].size
    g, a0, st2, _, rc2 = xo.dopri5_mlp_adjoint(om, t, sol, gy, nthreads=nthreads)
    dt = time.perf_counter() - t0
    assert rc == 0 and rc2 == 0
    return int(st.n_attempts.sum() + st2.n_attempts.sum()), dt


def cpu_baseline(target_s=12.0):
    from oracle import xde_oracle as xo

    xo.build()
    cores = os.cpu_count() or 1
    w, y0, t = workload(1 << 12)
    om = xo.MLP(*w, pre="cube")
    n, dt = oracle_pass(xo, om, y0, t, cores)           # calibration (also warms the threads)
    Bs = int(min(1 << 20, max(1 << 12, (1 << 12) * target_s / max(dt, 1e-3))))
    Bs = 1 << int(np.floor(np.log2(Bs)))
    w, y0, t = workload(Bs)
    n, dt = oracle_pass(xo, om, y0, t, cores)
    try:  # cfg1 (BASELINE configs[0]: the reference's own CPU-runnable case): one ode_demo call, batch 20, on one core
        w1, y1, t1 = workload(20)
        t0 = time.perf_counter()
        for _ in range(20):
            xo.dopri5_mlp(om, y1, t1, controller="batch", nthreads=1)
        cfg1_ms = (time.perf_counter() - t0) / 20 * 1e3
    except Exception:
        cfg1_ms = None
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port", "cfg1_call_ms_one_core": cfg1_ms,
            "sample": f"first {Bs} trajectories of the cfg2 batch (2^20), forward+adjoint, OpenMP over {cores} threads, "
                      f"{dt:.1f} s; Paddle CPU reference not installable offline -> C oracle port of the reference algorithm"}


def run_reference(args):
    """Reference arm: the reference's algorithm on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import xde_oracle as xo

    xo.build()
    cores = os.cpu_count() or 1
    Bs = args.ref_batch
    w, y0, t = workload(Bs)
    om = xo.MLP(*w, pre="cube")
    for _ in range(args.warmup):
        oracle_pass(xo, om, y0, t, cores)
    tot_n, tot_t = 0, 0.0
    for _ in range(args.steps):
        n, dt = oracle_pass(xo, om, y0, t, cores)
        tot_n += n
        tot_t += dt
    v = tot_n / tot_t
    sample = (f"each step = forward+adjoint over the first {Bs} trajectories of the cfg2 batch, "
              f"OpenMP over {cores} threads (C oracle port; Paddle not installable offline)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the B200 arm's config (the contract: "on your arm's config"); what one reference step really integrates is the
        # bounded sample named in cpu_baseline.sample / reference_sample -- the rate is per trajectory-step
        "config": config_dict(args.batch // max(args.gpus, 1) if args.scaling == "strong" else args.batch,
                              max(args.gpus, 1), "device", args),
        "reference_sample": {"trajectories_per_step": Bs, "of": args.batch, "where": "host cores, no GPU"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def config_dict(B, n, where, args=None):
    strong = args is not None and args.scaling == "strong"
    return {"workload": "cfg2: odeint_adjoint dopri5 rtol=1e-7 atol=1e-9, MLP 2-50-2 (y**3), "
                        "t=linspace(0,25,1000)[:10], loss=mean|y_T|",
            "batch_per_gpu": B, "global_batch": (args.batch if strong else B * n), "state_dim": 2, "hidden": 50,
            "n_out_times": 10,
            "controller": "trajectory", "adjoint_norm": "seminorm", "parallelism": f"batch-sharded x{n}",
            "l2": "256 MiB buffer written between steps (inside the timed region); working set 168 MiB > 126 MB L2"
                  if where == "device" else "n/a"}


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region.  NVML from a thread every ~20 ms
    (the timed region of the default run is ~0.1 s: `nvidia-smi -lms` takes longer than that to print its
    first row; every 4 ms the NVML queries made the host side of a launch several times slower -- they
    contend with the CUDA driver -- which showed as 2-5 ms bubbles between short kernels, r2c);
    `nvidia-smi` is the fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_id):
        self.rows, self.p, self.thr, self.stop_flag = [], None, None, False
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByUUID(gpu_id.encode() if isinstance(gpu_id, str) else gpu_id) \
                if str(gpu_id).startswith("GPU-") else N.nvmlDeviceGetHandleByIndex(int(gpu_id))
            mx = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            reasons_fn = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                N.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self.stop_flag:
                    try:
                        r = int(reasons_fn(h))
                        self.rows.append((float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), mx,
                                          N.nvmlDeviceGetPowerUsage(h) / 1000.0,
                                          [k for k, b in bits.items() if r & b]))
                    except Exception:
                        pass
                    time.sleep(0.02)

            import threading
            self.thr = threading.Thread(target=loop, daemon=True)
            self.thr.start()
            return
        except Exception:
            self.thr = None
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_id), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.thr is not None:
            self.stop_flag = True
            self.thr.join(timeout=2)
            if not self.rows:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
            sm = [r[0] for r in self.rows]
            return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(self.rows[0][1]),
                    "power_w_max": float(max(r[2] for r in self.rows)), "samples": len(sm),
                    "reasons": sorted({k for r in self.rows for k in r[3]}), "source": "nvml"}
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.strip().lower() == "active":
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------
# secondary configs (not the headline): BASELINE.json configs[2] (cfg3) and [3] (cfg4), device-resident
# ---------------------------------------------------------------------------------------------------
def secondary_configs(ffma_tflops=None):
    """Kernel timings of the large-state fixed-grid paths, tensor-core (tcgen05) and FP32, with the
    roofline fractions of SURVEY 8(d).  Reported next to the headline, never instead of it."""
    try:
        import torch

        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_configs as bc
        if ffma_tflops:  # the FP32 fractions of the secondary entries use THIS run's probe, not a pasted constant
            bc.FFMA = float(ffma_tflops)
        res = []
        t_start = time.perf_counter()
        try:  # cfg1 (BASELINE configs[0]) is a 20-trajectory call: its latency, reported beside the throughput configs
            res.append(bc.cfg1_latency())
        except Exception as e:
            res.append({"config": "cfg1 latency", "error": f"{type(e).__name__}: {e}"})
        for fn, kw in ((bc.cfg2_batch, {"norm": "mixed"}), (bc.cfg3, {"math": "tensor"}), (bc.cfg3, {"math": "fp32"}),
                       (bc.cfg4, {"math": "tensor", "B": 1 << 22}), (bc.cfg4, {"math": "fp32"}),
                       (bc.cfg4, {"math": "tensor", "generated": True, "B": 1 << 22}), (bc.cfg5, {}),
                       (bc.cfg5_real_size, {})):
            if time.perf_counter() - t_start > 150.0:  # the default run must end within minutes
                res.append({"config": f"{fn.__name__} {kw}", "skipped": "secondary time budget (150 s) spent"})
                continue
            try:  # one failing configuration must not take the others (or the headline) with it
                res.append(fn(**kw))
            except Exception as e:
                res.append({"config": f"{fn.__name__} {kw}", "error": f"{type(e).__name__}: {e}"})
            torch.cuda.empty_cache()
        return res
    except Exception as e:  # never lose the headline line over a secondary measurement
        return {"error": f"{type(e).__name__}: {e}"}


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    import paddlexde_b200 as px
    from paddlexde_b200 import _lib, _tensor as T
    from paddlexde_b200.functional.odeint_adjoint import adjoint_backward

    from paddlexde_b200 import distributed as pxd

    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU path"
    rank, world, local = pxd.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    reduce_grads = pxd.grad_allreduce()
    lib = _lib.lib()
    if args.scaling == "strong":  # the 1M-trajectory batch of the north star, split across the ranks
        lo, hi = pxd.shard_rows(args.batch, rank, world)
        w, y0_all, t = workload(args.batch, seed=0)
        y0_np = np.ascontiguousarray(y0_all[lo:hi])
        B = hi - lo
    else:
        B = args.batch
        w, y0_np, t = workload(B, seed=rank)
    field = px.MLPField(*w, pre="cube")
    y0 = torch.from_numpy(y0_np).to(dev)
    xde = px.xde.BaseODE(field, y0, t)
    gy = torch.zeros((t.size, B, 2), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    inv_n = 1.0 / ((args.batch if args.scaling == "strong" else B) * 2)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    k_ev = {"fwd": [], "adj": []}

    def step(timed):
        flush.zero_()
        s = px.Dopri5(xde=xde, y0=y0, rtol=1e-7, atol=1e-9, controller="trajectory", check_status=False)
        e = [ev() for _ in range(4)] if timed else None
        if timed:
            e[0].record()
        sol = s.integrate(t)
        if timed:
            e[1].record()
        torch.sign(sol[-1], out=gy[-1])
        gy[-1].mul_(inv_n)
        if timed:
            e[2].record()
        g, _, st, _ = adjoint_backward(field, t, sol, gy, check_status=False)
        if timed:
            e[3].record()
            k_ev["fwd"].append((e[0], e[1]))
            k_ev["adj"].append((e[2], e[3]))
        reduce_grads(g)  # the only collective on the path: 252 floats over NCCL/NVLink (no-op at N = 1)
        return s, st, g

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        s, st, g = step(False)
    barrier()
    fwd_stats, adj_stats = s.read_stats(), st.read()
    assert fwd_stats.status == 0 and adj_stats.status == 0, "solver status != OK"
    n_traj_steps = fwd_stats.n_attempts + adj_stats.n_attempts

    uuid = None
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        uuid = local
    sampler = ClockSampler(uuid) if rank == 0 else None
    n0 = px.launch_count()
    barrier()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.steps):
        s, st, g = step(True)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = px.launch_count() - n0
    clocks = sampler.stop() if sampler else None
    a2 = st.read()
    assert a2.n_attempts == adj_stats.n_attempts and s.read_stats().n_attempts == fwd_stats.n_attempts
    ms_fwd = float(np.mean([a.elapsed_time(b) for a, b in k_ev["fwd"]]))
    ms_adj = float(np.mean([a.elapsed_time(b) for a, b in k_ev["adj"]]))

    # ---- end to end through the public API, host buffers ----
    tw = [torch.tensor(a, device=dev, requires_grad=True) for a in w]
    field_e = px.MLPField(*tw, pre="cube")
    y0_pin = torch.from_numpy(y0_np).pin_memory()
    t_host = torch.from_numpy(t)
    # input pipeline: every step's y0 is copied host -> device inside the timed region, on a copy stream, into the
    # buffer the previous step is not using, so the copy of step n+1 overlaps the solve of step n
    copy_stream = torch.cuda.Stream(device=dev)
    y0_bufs = [torch.empty_like(y0), torch.empty_like(y0)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]

    def h2d(i):
        with torch.cuda.stream(copy_stream):
            y0_bufs[i % 2].copy_(y0_pin, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_step(i, last, opts):
        flush.zero_()
        for p in tw:
            p.grad = None
        torch.cuda.current_stream().wait_event(ready[i % 2])
        if not last:
            h2d(i + 1)  # the other buffer: its last reader (step i-1) finished before that step's loss.item()
        sol = px.odeint_adjoint(field_e, y0_bufs[i % 2], t_host, solver=px.Dopri5,
                                options={"controller": "trajectory", **opts.get("options", {})})
        loss = sol[-1].abs().sum() * inv_n   # mean|y_T| over the GLOBAL batch
        loss.backward()
        flat = reduce_grads(torch.cat([p.grad.reshape(-1) for p in tw]))
        return float(loss.item()), flat.cpu()

    def e2e_run(n, opts):
        h2d(0)
        for i in range(n):
            out = e2e_step(i, i == n - 1, opts)
        return out

    def e2e_time(opts):
        e2e_run(5, opts)  # the first backward passes through autograd grow the allocator pools
        barrier()
        e0, e1 = ev(), ev()
        e0.record()
        out = e2e_run(args.steps, opts)
        e1.record()
        barrier()
        return e0.elapsed_time(e1), out

    ms_e2e, (loss_v, g_host) = e2e_time({})  # the default call: the headline e2e
    # the same with the forward solve's assertions raised by the call itself, as the reference does (the host waits
    # between the two solves); reported next to the headline
    ms_e2e_sync, _ = e2e_time({"options": {"check_status": True}})

    # ---- the same training step captured ONCE in a CUDA graph and replayed (H2D of y0, forward, loss, backward,
    # gradient all-reduce, D2H of loss + gradients + both status words are all nodes of the graph): what a training loop
    # at small per-GPU batches (strong scaling) should do -- the eager step pays ~0.4 ms of host work per step
    ms_e2e_graph, graph_note = None, None
    try:
        if world > 1 and not args.graph_e2e:
            # measured once at N = 2 (r2i: 3.56e9 trajectory-steps/s, gradients equal to the eager step's) -- but that run
            # then hung in the process-group teardown with the NCCL all-reduce captured in the graph, so a default
            # multi-rank run leaves the capture out (--graph-e2e forces it)
            raise RuntimeError("skipped at N > 1 (pass --graph-e2e)")
        n_p = sum(p.numel() for p in tw)
        out_pin = torch.empty(n_p + 1, dtype=torch.float32).pin_memory()
        st_pin = torch.empty(8, dtype=torch.int64).pin_memory()
        y0_g = torch.empty_like(y0)
        gopts = {"controller": "trajectory", "check_status": False}

        def graph_body():
            y0_g.copy_(y0_pin, non_blocking=True)
            for p in tw:
                p.grad = None
            sol = px.odeint_adjoint(field_e, y0_g, t_host, solver=px.Dopri5, options=gopts)
            loss = sol[-1].abs().sum() * inv_n
            loss.backward()
            flat = reduce_grads(torch.cat([p.grad.reshape(-1) for p in tw] + [loss.detach().reshape(1)]))
            out_pin.copy_(flat, non_blocking=True)
            st_pin.copy_(px.odeint_adjoint.last["stats_pair"].both, non_blocking=True)

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                graph_body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            graph_body()

        def graph_step():
            flush.zero_()
            graph.replay()
            torch.cuda.current_stream().synchronize()
            return float(out_pin[-1]), out_pin[:-1].clone()

        for _ in range(3):
            lg, gg = graph_step()
        barrier()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(args.steps):
            lg, gg = graph_step()
        e1.record()
        barrier()
        ms_e2e_graph = e0.elapsed_time(e1)
        ok = (int(st_pin[3]) & 0xffffffff) == 0 and (int(st_pin[7]) & 0xffffffff) == 0
        same = bool(torch.allclose(gg, g_host, rtol=1e-5, atol=1e-6 * float(g_host.abs().max())))
        graph_note = f"status words OK={ok}, gradients equal to the eager step's within rtol 1e-5: {same}"
    except Exception as e:  # reported, never fatal: the eager number is the headline
        graph_note = str(e) if "skipped" in str(e) else f"capture failed: {type(e).__name__}: {e}"

    # ---- FP32 pipe ceiling (measured) ----
    sink = torch.zeros(1, device=dev)
    nfl = C.c_int64(0)
    for _ in range(2):
        _lib.check(lib.xde_probe_ffma_f32(1 << 14, T.ptr(sink), C.byref(nfl), T.stream()))
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    _lib.check(lib.xde_probe_ffma_f32(1 << 16, T.ptr(sink), C.byref(nfl), T.stream()))
    e1.record()
    torch.cuda.synchronize()
    ffma_tflops = nfl.value / (e0.elapsed_time(e1) * 1e-3) / 1e12

    ms_cfg3 = 0.0
    if world > 1 and not args.no_secondary:
        # BASELINE config 3 is "8xB200 batch-sharded": at N > 1 every rank also integrates ITS 2^17-trajectory share of the
        # cfg3 ensemble (RK4, 64-256-64, 100 steps, tcgen05) -- no communication on that path; the share's kernel time
        # joins the max-over-ranks reduction below (no extra collective).  A failure is reported as 0, never raised:
        # an exception on one rank would leave the others waiting in the reduction.
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_configs as bc
            ms_cfg3 = float(bc.cfg3(B=args.cfg3_share, math="tensor")["ms"])
        except Exception:
            ms_cfg3 = 0.0
    if world > 1:
        tt = torch.tensor([ms, ms_e2e, ms_fwd, ms_adj, ms_e2e_sync, ms_e2e_graph or 0.0, ms_cfg3,
                           0.0 if ms_cfg3 > 0 else 1.0], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_fwd, ms_adj, ms_e2e_sync, mg, ms_cfg3, cfg3_missing = tt.tolist()
        ms_e2e_graph = mg if ms_e2e_graph else None
        if cfg3_missing > 0:  # some rank could not measure its share: report nothing rather than a partial maximum
            ms_cfg3 = 0.0
        cnt = torch.tensor([n_traj_steps, fwd_stats.n_attempts, adj_stats.n_attempts], device=dev, dtype=torch.int64)
        cnt_local = cnt.clone()
        dist.all_reduce(cnt)
        total_steps = int(cnt[0])
    else:
        total_steps = n_traj_steps
    if rank == 0:
        hbm, which = peaks()
        value = total_steps * args.steps / (ms * 1e-3)
        dom = "adj" if ms_adj >= ms_fwd else "fwd"
        k_name = "dopri5_adj_kernel" if dom == "adj" else "dopri5_fwd_small_kernel"
        k_attempts = adj_stats.n_attempts if dom == "adj" else fwd_stats.n_attempts
        k_nfe = adj_stats.nfe if dom == "adj" else fwd_stats.nfe
        k_ms = ms_adj if dom == "adj" else ms_fwd
        k_bytes = (BYTES_ADJ if dom == "adj" else BYTES_FWD) * k_attempts
        k_flops = (FLOPS_ADJ if dom == "adj" else FLOPS_FWD) * k_attempts
        achieved_gbs = k_bytes / (k_ms * 1e-3) / 1e9
        achieved_tf = k_flops / (k_ms * 1e-3) / 1e12
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        mhz = (clocks or {}).get("sm_mhz") or 1965.0
        # FMA-pipe model of the adjoint kernel, evaluated on THIS run's counters: per field+VJP evaluation of one
        # trajectory 25 hidden-unit pairs x 27 packed instructions (2 x 3 first layer, 16 rational tanh, 8 VJP / second
        # layer) + 5 packed fold instructions per pair and column, each holding one scheduler's FMA pipe for 2 cycles
        # per warp of 32 trajectories (tools/probe_issue.cu).  Evaluations really executed: 2 per segment start
        # (f0 + probe; the reference counts 3) + 6 per attempt.
        n_seg = (t.size - 1) * B
        evals = 6 * adj_stats.n_attempts + 2 * n_seg
        cyc = 2.0 * (25 * 27 + 32 * 5) / 32.0
        bound_ms = evals * cyc / (sms * 4 * mhz * 1e3)
        prof = latest_profile(k_name, k_ms)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(B, world, "device", args),
            "trajectory_steps_per_step": {"forward": fwd_stats.n_attempts, "adjoint": adj_stats.n_attempts,
                                          "accepted_forward": fwd_stats.n_accepted, "accepted_adjoint": adj_stats.n_accepted,
                                          "nfe_forward": fwd_stats.nfe, "nfe_adjoint": adj_stats.nfe, "per": "rank 0"},
            "kernel_ms": {"dopri5_fwd_small_kernel": ms_fwd, "dopri5_adj_kernel(+cast)": ms_adj},
            # D = 2, H = 50: 75-112 FLOP/B of scalar FP32 + 50 tanh per evaluation, K = 2 / N = 2 GEMVs (degenerate for
            # an MMA): the FP32 FMA pipe binds, not HBM.  peak = the FFMA probe of THIS run.
            "roofline": {"bound": "fp32", "kernel": k_name, "achieved": achieved_tf, "peak": ffma_tflops,
                         "unit": "TFLOP/s", "frac": achieved_tf / ffma_tflops,
                         "peak_source": "xde_probe_ffma_f32 timed in this run (8 independent FFMA chains per thread, "
                                        "every SM); not in MEASURED_PEAKS.json, which holds HBM and bf16 only",
                         "algorithmic_flops_per_launch": k_flops,
                         "traffic": prof["traffic"] if prof and prof["matches_this_run"] else None,
                         "traffic_source": (prof or {}).get("source") if prof and prof["matches_this_run"] else
                         "no committed ncu capture matches this run's kernel time within 2 %",
                         "note": "algorithmic FLOPs = SURVEY 8(d)'s 72*D*H (adjoint) / 24*D*H (forward) per trajectory-step: "
                                 "the two GEMVs and their VJPs only.  The arithmetic specification (rational 13/6 tanh + IEEE "
                                 "division, bit-exact against the oracle) adds 16 of the 27 packed FMA-pipe instructions per "
                                 "hidden-unit pair that this count leaves out: see fma_pipe_model"},
            "roofline_hbm": {"bound": "hbm", "kernel": k_name, "achieved": achieved_gbs, "peak": hbm, "unit": "GB/s",
                             "frac": achieved_gbs / hbm, "peak_source": which, "algorithmic_bytes_per_launch": k_bytes,
                             "note": "reported because the contract asks for it; the kernel keeps (y, a) in registers across "
                                     "the attempts of a segment, so even the algorithmic bytes are not moved"},
            "fma_pipe_model": {"evaluations": evals, "fma_pipe_cycles_per_evaluation_and_warp": cyc * 32.0,
                               "bound_ms": bound_ms, "frac_of_bound": bound_ms / ms_adj, "sm_mhz": mhz,
                               "note": "time the adjoint launch would take if the packed FP32 instructions the arithmetic "
                                       "specification requires (field + VJP + parameter-gradient fold) kept every scheduler's "
                                       "FMA pipe busy at the sampled SM clock; controller arithmetic not counted"},
            "ncu": prof,
            "e2e": {"value": total_steps * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(y0_np.nbytes + t.nbytes), "d2h_bytes_per_step": int(g_host.numel() * 4 + 4 + 64),
                    "ms_per_step": ms_e2e / args.steps,
                    "api": "paddlexde_b200.odeint_adjoint(field, y0, t, solver=Dopri5); loss.backward()  (a training step: the "
                           "forward solve's assertions are raised by backward(), one D2H for both status words)",
                    "with_synchronous_status_check": {
                        "value": total_steps * args.steps / (ms_e2e_sync * 1e-3), "ms_per_step": ms_e2e_sync / args.steps,
                        "api": "odeint_adjoint(..., options={'check_status': True}): the call itself raises, as the reference "
                               "does (the host waits between the two solves)"},
                    "with_cuda_graph": ({"value": total_steps * args.steps / (ms_e2e_graph * 1e-3),
                                         "ms_per_step": ms_e2e_graph / args.steps} if ms_e2e_graph else {}) | {
                        "api": "the same calls captured once with torch.cuda.graph (options={'check_status': False}; status "
                               "words copied to pinned memory by the graph and checked after each replay) and replayed",
                        "note": graph_note},
                    "input_pipeline": "one H2D copy of y0 per step from pinned memory on a copy stream, double-buffered: "
                                      "the copy for step n+1 overlaps the solve of step n (all inside the timed region)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world > 1 and ms_cfg3 > 0:
            out["cfg3_sharded"] = {
                "config": f"cfg3: RK4 3/8 rule, MLP 64-256-64, 100 steps, {args.cfg3_share} trajectories per GPU x {world} GPUs "
                          f"(global {world * args.cfg3_share}; BASELINE: 2^20 on 8 GPUs), tcgen05 field, batch-sharded, "
                          "no collective",
                "ms_max_over_ranks": ms_cfg3, "traj_steps_per_s": world * args.cfg3_share * 100 / (ms_cfg3 * 1e-3),
                "timing": "CUDA events around the solve on each rank (median of 5 after 2 warm-up solves, device-resident "
                          "inputs), max over ranks through the same all-reduce as the headline timings"}
        if not args.no_cpu and world == 1:  # the CPU leg is timed at N = 1 only (rank 0)
            out["cpu_baseline"] = cpu_baseline(args.cpu_seconds)
        if world == 1 and not args.no_secondary:
            out["secondary"] = secondary_configs(ffma_tflops)
        print(json.dumps(out), flush=True)
    if world > 1:
        # Every collective of the run is behind us (the last one is the all-reduce of the timings above).  Leave without
        # the process-group teardown: r2i showed a rank blocking in it (barrier / destroy / interpreter exit) after the
        # JSON line had been printed, which cost the whole 8-GPU lease.  torchrun sees exit code 0 from every rank.
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1 << 20, help="trajectories per GPU")
    ap.add_argument("--ref-batch", type=int, default=1 << 14, help="trajectories per reference-arm step")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the cfg3/cfg4 kernel timings")
    ap.add_argument("--cfg3-share", type=int, default=1 << 17,
                    help="trajectories per GPU of the cfg3 ensemble share measured at N > 1 (BASELINE: 2^20 over 8 GPUs)")
    ap.add_argument("--graph-e2e", action="store_true", help="capture the e2e step in a CUDA graph at N > 1 as well")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch trajectories per GPU; strong: --batch trajectories in total, split across the GPUs")
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5"],
                    help="BASELINE.json config to measure (cfg2 = the headline; the others print the same line schema)")
    args = ap.parse_args()
    if args.config != "cfg2":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_lines

        return bench_lines.main(args, ClockSampler, peaks)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
