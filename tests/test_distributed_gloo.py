"""World-size-2 `gloo` test of the N > 1 host path: batch sharding + the adjoint
parameter-gradient all-reduce.  There is no GPU here, so each rank's partial gradient comes from the
CPU oracle standing in for the adjoint kernel (the checker role it has everywhere in tests/)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, B, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from oracle import xde_oracle as xo
    from paddlexde_b200 import distributed as D
    from tests.problems import cfg2_tspan, cfg2_y0, spiral_weights

    r, w, _ = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    om = xo.MLP(*spiral_weights(), pre="cube")
    y0, t = cfg2_y0(B), cfg2_tspan(5)
    lo, hi = D.shard_rows(B, rank, world)
    sol, _, _, rc = xo.dopri5_mlp(om, y0[lo:hi], t)
    gy = np.zeros_like(sol)
    gy[-1] = np.sign(sol[-1]) / (B * 2)              # loss = mean|y_T| over the GLOBAL batch
    g, _, st, _, rc2 = xo.dopri5_mlp_adjoint(om, t, sol, gy)
    assert rc == 0 and rc2 == 0
    gt = torch.from_numpy(g.copy())
    D.grad_allreduce()(gt)
    cnt = torch.tensor([int(st.n_attempts.sum())])
    torch.distributed.all_reduce(cnt)
    np.save(os.path.join(out_dir, f"g{rank}.npy"), gt.numpy())
    np.save(os.path.join(out_dir, f"n{rank}.npy"), cnt.numpy())
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_shard_rows_partition():
    from paddlexde_b200.distributed import shard_rows

    for n in (0, 1, 7, 8, 1 << 20):
        for world in (1, 2, 3, 8):
            cuts = [shard_rows(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_rows(4, 2, 2)


def test_sde_shard_options_reproduce_the_global_increment_stream():
    """The counter-based Brownian generator is addressed by the global trajectory index: the increments of
    every shard (NumPy restatement of the device generator) tile the unsharded table exactly."""
    from oracle.philox_np import brownian_increments
    from paddlexde_b200.distributed import sde_shard_options, shard_rows

    t = np.linspace(0, 1, 4).astype(np.float32)
    whole = brownian_increments(99, t, 37, 5)
    for world in (2, 3):
        parts = []
        for r in range(world):
            lo, hi = shard_rows(37, r, world)
            o = sde_shard_options(99, 37, r, world)
            assert o == {"bm_seed": 99, "bm_offset": lo}
            parts.append(brownian_increments(o["bm_seed"], t, hi - lo, 5, offset=o["bm_offset"]))
        assert np.array_equal(np.concatenate(parts, axis=1), whole)


@pytest.mark.timeout(300)
def test_two_rank_gradient_allreduce_equals_single_process(tmp_path, oracle):
    from tests.problems import cfg2_tspan, cfg2_y0, spiral_weights

    B, world = 101, 2  # odd: ragged shards
    mp.spawn(_worker, args=(world, _free_port(), B, str(tmp_path)), nprocs=world, join=True)
    g0, g1 = np.load(tmp_path / "g0.npy"), np.load(tmp_path / "g1.npy")
    assert np.array_equal(g0, g1), "every rank must hold the same reduced gradient"
    om = oracle.MLP(*spiral_weights(), pre="cube")
    y0, t = cfg2_y0(B), cfg2_tspan(5)
    sol, _, _, _ = oracle.dopri5_mlp(om, y0, t)
    gy = np.zeros_like(sol)
    gy[-1] = np.sign(sol[-1]) / (B * 2)
    g, _, st, _, _ = oracle.dopri5_mlp_adjoint(om, t, sol, gy)
    # one controller per trajectory: sharding changes nothing but the summation order of the gradients
    np.testing.assert_allclose(g0, g, rtol=1e-5, atol=1e-6 * np.abs(g).max())
    assert int(np.load(tmp_path / "n0.npy")[0]) == int(st.n_attempts.sum())
