"""Multi-GPU plumbing: one process per GPU, imagesets sharded over ranks, ONE NCCL all-reduce
of the partial normal equations per H/b build.

torch.distributed is used only for rendezvous (broadcasting the NCCL unique id) and for the
max-over-ranks timing in bench.py; the data-path collective is issued by ``libb200ba.so`` on
its own communicator and stream.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np

from .api import BundleAdjuster, nccl_unique_id
from .cabi import FlatProblem, FlatState


def env_rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_problem(problem: FlatProblem, rank: int, world: int) -> Tuple[FlatProblem, np.ndarray]:
    """Observations owned by ``rank``: imageset i belongs to rank i % world (all cameras of an
    imageset stay together, so pose blocks are produced on one rank only). Returns the shard and
    the indices of its observations in the full problem (for last_projection scatter/gather)."""
    idx = problem.shard_indices(rank, world)
    return problem.shard(rank, world), idx


def shard_state(state: FlatState, idx: np.ndarray) -> FlatState:
    """State for a shard: parameters are replicated, last_projection is restricted."""
    st = state.copy()
    if st.last_projection is not None:
        st.last_projection = np.ascontiguousarray(st.last_projection[idx])
    return st


def init_process_group(backend: Optional[str] = None):
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return dist
    rank, world, local = env_rank_world()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return dist


def make_sharded_adjuster(problem: FlatProblem, device: int = -1) -> Tuple[BundleAdjuster, np.ndarray]:
    """Creates this rank's BundleAdjuster on its shard and joins the NCCL communicator."""
    import torch.distributed as dist
    rank, world, local = env_rank_world()
    if world == 1:
        return BundleAdjuster(problem, device), np.arange(problem.n_obs)
    shard, idx = shard_problem(problem, rank, world)
    adj = BundleAdjuster(shard, local if device < 0 else device)
    box = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    adj.comm_init(box[0], rank, world)
    return adj, idx
