"""Batch sharding across GPUs.

Trajectories are independent given the field parameters, so the path shards by batch: one process per
GPU, contiguous rows of ``y0`` (and of ``dW[:, rows]``, ``his[rows]``) per rank, parameters replicated.
Forward / SDE / gather need no communication.  The ONLY collective is one all-reduce (sum) of the
adjoint parameter-gradient vector per backward -- the counterpart of the DataParallel gradient
all-reduce in the reference's example trainer (example/D3STN/train_dde.py:201-202,454-456)."""
from __future__ import annotations

import ctypes as C
import os

from . import _tensor as T
from ._lib import check, lib


def shard_rows(n: int, rank: int, world: int):
    """[start, stop) of the contiguous rows owned by `rank`; the first n % world ranks get one extra row."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def sde_shard_options(seed: int, n: int, rank: int, world: int) -> dict:
    """`options=` for `sdeint` on this rank's rows when the Brownian increments are generated on the device:
    the generator is addressed by the GLOBAL trajectory index, so passing the shard's first row as
    `bm_offset` makes the sharded run reproduce the unsharded one bit for bit (no 8 GiB table to scatter)."""
    lo, _ = shard_rows(n, rank, world)
    return {"bm_seed": int(seed), "bm_offset": lo}


def init_from_env(backend: str | None = None):
    """One process per GPU, launched by torchrun: returns (rank, world, local_rank)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        return rank, world, local
    import torch
    import torch.distributed as dist

    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def grad_allreduce(group=None):
    """The hook for ``options={"grad_allreduce": ...}`` of odeint_adjoint: sums the flat parameter-gradient
    vector over the ranks in place (NCCL over NVLink on GPUs; a no-op for a single process)."""

    def hook(g):
        if T.torch is None:
            return g
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
        return g

    return hook


def nccl_grad_allreduce(comm: int):
    """The same hook WITHOUT torch.distributed: `comm` is the address of an ncclComm_t the caller created (Paddle's,
    its own, ...); the sum runs through the library's xde_allreduce_grads on the stream of the gradient buffer
    (include/xde_b200.h; SURVEY 8(b) proposed exactly this entry)."""

    def hook(g):
        check(lib().xde_allreduce_grads(C.c_void_p(int(comm)), T.ptr(g), g.numel(), T.stream(g)))
        return g

    return hook
