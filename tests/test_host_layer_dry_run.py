"""The Python host layer, executed on the CPU: the GPU-marked test files run in a subprocess with two test doubles
(tests/host_dry_run.py): the C ABI is answered by the CPU oracle on the same raw pointers and CUDA placement is sent to
the CPU.  This checks what lives ABOVE the C ABI -- argument marshalling against include/xde_b200.h, layouts, option
routing, the autograd adapters, status / attempt-log decoding -- and that the GPU tests themselves are executable.  It
says nothing about the kernels (both sides are the oracle's numbers by construction); the parity tests proper are the
same files under `-m gpu` on a B200.  Deselected: tests that measure the tensor-core entries' own accuracy (their double
is the FP32 oracle), tests of shape refusals / status words raised by the kernels, and the full-size cases."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SKIP = ("tensor", "full_size", "status", "test_midpoint_fixed_solver", "tiled_sde_cfg4", "baseline_size")
DESELECT = ("tests/test_gpu_parity.py::test_other_tableaux_large_state[16-64-square-33-Bosh3-1e-05]",)  # ends with a refusal


def test_gpu_test_files_run_against_the_host_layer_doubles():
    # the oracle's OpenMP team spins between the many short calls: a few threads are much faster than all cores here
    env = dict(os.environ, XDE_DRY_RUN="1", PYTHONPATH=ROOT, OMP_NUM_THREADS="4", OMP_WAIT_POLICY="passive")
    cmd = [sys.executable, "-m", "pytest", "tests/test_zz_reference_run_gpu.py", "tests/test_gpu_round2.py",
           "tests/test_gpu_parity.py", "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider",
           "-k", " and ".join(f"not {s}" for s in SKIP)] + [f"--deselect={d}" for d in DESELECT]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = "\n".join(r.stdout.splitlines()[-25:])
    assert r.returncode == 0, tail + "\n" + r.stderr[-2000:]
    assert " passed" in tail and "failed" not in tail, tail


def _run(code, timeout=900):
    env = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS="4", OMP_WAIT_POLICY="passive")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + "\n" + r.stderr[-3000:]
    return r.stdout


def test_smoke_runs_against_the_host_layer_doubles():
    out = _run("from tests import host_dry_run; host_dry_run.install()\n"
               "import __graft_entry__ as g; g.smoke()")
    assert "smoke ok" in out


def test_bench_line_schema_against_the_host_layer_doubles():
    """bench.py end to end on the doubles (a small batch, no secondary configs): the JSON line carries every key the
    contract names.  The numbers are meaningless here (host wall clock around the CPU oracle)."""
    import json

    out = _run("import sys\n"
               "from tests import host_dry_run; host_dry_run.install()\n"
               "sys.argv = ['bench.py', '--steps', '2', '--warmup', '1', '--batch', '2048', '--no-secondary', '--cpu-seconds', '0.3']\n"
               "import bench; bench.main()")
    line = json.loads([ln for ln in out.splitlines() if ln.startswith("{")][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["config"]["workload"].startswith("cfg2") and line["dtype"] == "f32" and line["vs_baseline"] is None
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(line["cpu_baseline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["gpu_launches"] > 0 and line["steps"] == 2


def _free_port():
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_bench_two_ranks_under_torchrun_on_gloo():
    """The driver's N > 1 launch (`python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N`) on the
    doubles with `gloo`: both ranks leave with exit code 0 (the round-2 exit path: no process-group teardown), rank 0
    prints ONE line, weak scaling doubles the global batch and strong scaling splits it."""
    import json

    for scaling, per_gpu, total in (("weak", 512, 1024), ("strong", 256, 512)):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
               "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_dry_run_bench.py"),
               "--gpus", "2", "--steps", "1", "--warmup", "1", "--batch", "512", "--scaling", scaling, "--cfg3-share", "16"]
        r = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-1500:] + "\n" + r.stderr[-3000:]
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        assert len(lines) == 1, lines
        line = json.loads(lines[0])
        assert line["n_gpus"] == 2 and line["scaling"] == scaling
        assert line["config"]["batch_per_gpu"] == per_gpu and line["config"]["global_batch"] == total
        assert "cpu_baseline" not in line  # rank 0 at N = 1 only
        assert line["cfg3_sharded"]["ms_max_over_ranks"] > 0  # BASELINE config 3's share per rank, max over ranks


def test_reference_arm_line_and_its_torchrun_launch():
    """`bench.py --impl reference` needs no GPU: rank 0 times the oracle port on the host cores and prints the line (the
    B200 arm's metric / unit / config, `impl`, `cpu_baseline`, zero-copy `e2e`), every other rank exits 0 without work."""
    import json

    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "1", "--ref-batch", "512"]
    r = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + "\n" + r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, lines
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "trajectory-steps/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["batch_per_gpu"] == 1 << 20 and line["config"]["parallelism"] == "batch-sharded x2"
    assert line["reference_sample"]["trajectories_per_step"] == 512


def test_training_example_runs_and_the_field_follows_the_optimizer():
    """examples/ode_demo.py (the reference's example/ode_demo.py flow through the public API) on the doubles: the loss is
    finite, every parameter moves, and the solves see the optimizer's in-place updates (MLPField re-reads the caller's
    tensors when their version counters change: without that the loss sequence would not react to the steps)."""
    out = _run("from tests import host_dry_run; host_dry_run.install()\n"
               "import runpy, numpy as np\n"
               "m = runpy.run_path('examples/ode_demo.py')\n"
               "a, pa = m['main'](['--steps', '12', '--lr', '0.01'])\n"
               "b, pb = m['main'](['--steps', '12', '--lr', '0.0'])\n"
               "assert all(np.isfinite(a)) and all(np.isfinite(b))\n"
               "assert a[0] == b[0] and a[1:] != b[1:], 'the solves ignore the optimizer steps'\n"
               "assert np.mean(a[-4:]) < np.mean(b[-4:]), (a, b)\n"
               "print('example ok', a[0], a[-1], b[-1])")
    assert "example ok" in out
