"""Multi-GPU plumbing: one process per GPU (torchrun), independent series sharded by
contiguous blocks, ONE all-reduce of the accumulated spectra.

Nothing here touches the data path between kernels: every rank transforms its own shard
with the same kernels; the only exchange is ``all_reduce(SUM)`` of ``[n_acc, M]`` doubles
(NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """(rank, world, local_rank); initialises torch.distributed when WORLD_SIZE > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keeps NCCL's version banner off stdout (CSV / JSON consumers)
        if backend is None:  # DYNPY_DIST_BACKEND=gloo: several ranks on one GPU (tests), CUDA tensors staged by gloo
            backend = os.environ.get("DYNPY_DIST_BACKEND") or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shutdown() -> None:
    """Tear the process group down (no-op for single-process runs)."""
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def rank_world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int | None = None, world: int | None = None) -> tuple[int, int]:
    """Contiguous, balanced block [lo, hi) of n independent units for this rank.
    Blocks are multiples of 2 units where possible so real-series pair packing stays aligned."""
    if rank is None or world is None:
        rank, world = rank_world()
    pairs = (n + 1) // 2
    lo = (pairs * rank) // world * 2
    hi = (pairs * (rank + 1)) // world * 2
    return min(lo, n), min(hi, n)


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place SUM over ranks (no-op for a single process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def h2d_sharded(host: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """Host -> device copy of an array every rank needs in full (the coordinates): each rank moves 1/world of
    dim 0 over ITS host link and the blocks are then exchanged GPU to GPU in ONE collective
    (``all_gather_into_tensor`` over NVLink when the blocks are equal, ``all_gather`` of views otherwise),
    instead of every rank pulling the whole array through PCIe.  Single process: a plain copy."""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    if world == 1 or host.shape[0] < world:
        out.copy_(host, non_blocking=True)
        return out
    rank = dist.get_rank()
    n = host.shape[0]
    cuts = [(n * r) // world for r in range(world + 1)]
    mine = out[cuts[rank]:cuts[rank + 1]]
    mine.copy_(host[cuts[rank]:cuts[rank + 1]], non_blocking=True)
    gather_blocks(out, cuts, mine)
    return out


def gather_blocks(full: torch.Tensor, cuts, mine: torch.Tensor) -> None:
    """Every rank holds block ``full[cuts[rank]:cuts[rank+1]]`` (= mine); afterwards every rank holds all of
    ``full``.  Equal blocks: ONE in-place all_gather_into_tensor; ragged blocks (the collective needs equal
    sizes): one broadcast per block."""
    world = dist.get_world_size()
    sizes = {cuts[r + 1] - cuts[r] for r in range(world)}
    if len(sizes) == 1 and full.is_contiguous():
        dist.all_gather_into_tensor(full, mine)     # in place: rank r's block is already at its final offset
    else:
        for r in range(world):
            dist.broadcast(full[cuts[r]:cuts[r + 1]], src=r)


def broadcast_object(obj, src: int = 0):
    """Rank `src`'s python object on every rank (single process: the object itself)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        box = [obj]
        dist.broadcast_object_list(box, src=src)
        return box[0]
    return obj


def all_gather_object(obj) -> list:
    """The python objects of all ranks, in rank order (single process: [obj])."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, obj)
        return parts
    return [obj]


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
